// instantiations of the lane-family kernels for N in [14, 15] (see lane_kernels.cuh)
#include "lane_kernels.cuh"
LANE_INSTANTIATE(14)
LANE_INSTANTIATE(15)
