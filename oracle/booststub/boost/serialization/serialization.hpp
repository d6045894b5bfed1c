// stand-in header (oracle/booststub, TEST INFRASTRUCTURE): just enough of boost::serialization for the reference's
// DBoW2 BowVector.h / FeatureVector.h to parse; their serialize() templates are never instantiated by the oracle.
#pragma once
namespace boost { namespace serialization {
class access {};
template <class Base, class Derived> Base& base_object(Derived& d) { return static_cast<Base&>(d); }
}}
