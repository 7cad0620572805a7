// nbk_emul.cpp -- TEST INFRASTRUCTURE: host build of one kernel program (see cuda_host_shim.h).
// Compiled once per (method, rhs) by tests/emul/build_emul.py with -DNBK_METHOD/-DNBK_RHS_T/-DNBK_TAG;
// exports emul_<tag>(kernel, args, n_threads): kernel 0 init, 1 run, 2 step.
#include "cuda_host_shim.h"
#include "nbk_program.cu"

#define EMUL_NAME NBK_CAT(emul_, NBK_TAG)

extern "C" int EMUL_NAME(int kernel, const nbk_kargs *a, long long n_threads) {
    blockDim.x = 1;
    threadIdx.x = 0;
    for (long long i = 0; i < n_threads; ++i) {
        blockIdx.x = (unsigned)i;
        if (kernel == 0) NBK_KERNEL(init)(*a);
        else if (kernel == 1) NBK_KERNEL(run)(*a);
        else if (kernel == 2) NBK_KERNEL(step)(*a);
#ifdef NBK_NEVENTS
        else if (kernel == 3) NBK_KERNEL(runev)(*a);
#endif
        else return -1;
    }
    return 0;
}
