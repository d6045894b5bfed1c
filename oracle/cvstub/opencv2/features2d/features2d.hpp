// stand-in header: see ../opencv.hpp (oracle/cvstub, TEST INFRASTRUCTURE)
#include "../opencv.hpp"
