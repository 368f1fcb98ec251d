// mh_fast.cu -- production instantiations of the fused step kernels (FMA contraction on).
#define MCGPU_NS fast
#include <stdlib.h>
#include "mh_kernels.cuh"
namespace mcgpu { namespace fast {
#include "mh_dispatch.inl"
}}
