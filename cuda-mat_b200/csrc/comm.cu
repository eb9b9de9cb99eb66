// comm.cu — multi-GPU plumbing (one process per GPU). NCCL is dlopen()ed on first use so that the
// single-GPU library has no link-time dependency on it.
#include "solver.h"
#include <dlfcn.h>

namespace cudamat {
struct Comm { int rank = 0, world = 1; };
}
using namespace cudamat;

extern "C" {
int cudamat_comm_unique_id(void *id128) {
    (void)id128;
    set_error("multi-GPU support is not compiled into this build yet");
    return CUDAMAT_E_COMM;
}
int cudamat_comm_init(cudamat_solver *s, const void *id128, int rank, int world) {
    (void)s; (void)id128; (void)rank; (void)world;
    set_error("multi-GPU support is not compiled into this build yet");
    return CUDAMAT_E_COMM;
}
}
