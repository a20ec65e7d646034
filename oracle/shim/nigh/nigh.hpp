#include "nigh_linear.hpp"
