// forwarding header: same include path as the reference tree
#pragma once
#include "sela_api.hpp"
