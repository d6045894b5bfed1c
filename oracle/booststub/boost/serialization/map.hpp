// stand-in header (oracle/booststub, TEST INFRASTRUCTURE)
#pragma once
#include "serialization.hpp"
