// vx_light.cu -- L-ext voxel light relaxation (placeholder translation unit; kernels land here).
#include "vx_kernels.hpp"
namespace vx {}
