// lna_eslice.cuh -- elliptical-slice kernels (filled in below)
#pragma once
#include "lna_kernels.cuh"
