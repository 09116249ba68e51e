// placeholder until the tcgen05 kernel lands: reports "shape not supported" so callers use the generic kernel
#include "common.cuh"
#include "kernels.h"
namespace rb {
int launch_dense_tc_f32(const DenseProblem<float>&, cudaStream_t) { return -1; }
}
