// instantiations of the lane-family kernels for N in [16] (see lane_kernels.cuh)
#include "lane_kernels.cuh"
LANE_INSTANTIATE(16)
