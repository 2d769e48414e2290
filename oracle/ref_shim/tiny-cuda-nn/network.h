#pragma once
// TEST INFRASTRUCTURE ONLY (oracle/_ref): envmap.cuh includes it, uses nothing from it on this path.
#include <tiny-cuda-nn/common.h>
