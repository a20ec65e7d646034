#include "nigh_shim.hpp"
