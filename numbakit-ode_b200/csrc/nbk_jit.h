// nbk_jit.h -- NVRTC program cache + CUDA driver-API loader (internal).
#ifndef NBK_JIT_H
#define NBK_JIT_H

#include <string>
#include <vector>

struct nbk_jit_request {
    int method;             // 23, 45, 1..5
    int n, n_params;
    std::string rhs_type;   // C++ type name of a built-in rhs (e.g. "nbk::rhs_decay<2,2>") or empty
    std::string user_src;   // user source defining nbk_rhs(...) / nbk_rhs_i(...) (used when rhs_type is empty)
    int cta_elems = 0;      // > 0: CTA-per-trajectory regime (nbk_program_cta.cu) with E components per thread
    int cta_maxt = 0;       // block-size cap the CTA kernels are compiled for
    int cta_full = 0;       // 1: n == block * E (every thread owns E valid components): no bounds checks compiled in
    int group = 0;          // CTA regime: G > 0 = sub-warp run / step kernels (nbk_group.cuh), G lanes per trajectory
    int n_events = 0;       // > 0: also build the run_events kernel with event_src's nbk_event(k, t, y, p) inlined
    std::string event_src;
};

struct nbk_jit_module {
    void *module = nullptr;  // CUmodule
    void *k_init = nullptr;  // CUfunction
    void *k_run = nullptr;
    void *k_step = nullptr;
    void *k_runev = nullptr;  // only in programs compiled with events
};

// Compile (or fetch from the in-process cache) the cubin of a request. Needs no GPU.
// Returns false and fills err (with the NVRTC log) on failure.
bool nbk_jit_compile(const nbk_jit_request &rq, std::vector<char> &cubin, std::string &err);

// Load a cubin into the current context and resolve the three kernels.
bool nbk_jit_load(const std::vector<char> &cubin, nbk_jit_module &mod, std::string &err);
// Drop one reference; modules nobody references stay loaded up to a small bound (least recently used are unloaded).
void nbk_jit_unload(nbk_jit_module &mod);
long nbk_jit_modules_loaded();

// cuLaunchKernel with a single by-value argument block.
bool nbk_jit_launch(void *func, unsigned grid, unsigned block, size_t smem, void *arg_block, void *stream,
                    std::string &err);
// opt in to more than 48 KB of dynamic shared memory / occupancy with dynamic shared memory
bool nbk_jit_set_smem(void *func, size_t smem, std::string &err);
bool nbk_jit_occupancy_smem(void *func, int block, size_t smem, int *blocks_per_sm, int *regs, std::string &err);
// cuOccupancyMaxActiveBlocksPerMultiprocessor / register count of a CUfunction.
bool nbk_jit_occupancy(void *func, int block, int *blocks_per_sm, int *regs, std::string &err);
// block size the kernel was compiled for (__launch_bounds__ -> CU_FUNC_ATTRIBUTE_MAX_THREADS_PER_BLOCK)
bool nbk_jit_block_size(void *func, int *block, std::string &err);

#endif
