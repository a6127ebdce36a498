// instantiations of the lane-family kernels for N in [1, 2, 3, 4, 5, 6] (see lane_kernels.cuh)
#include "lane_kernels.cuh"
LANE_INSTANTIATE(1)
LANE_INSTANTIATE(2)
LANE_INSTANTIATE(3)
LANE_INSTANTIATE(4)
LANE_INSTANTIATE(5)
LANE_INSTANTIATE(6)
