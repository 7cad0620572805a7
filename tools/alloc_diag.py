import os, time, torch
print("ALLOC_CONF", os.environ.get("PYTORCH_CUDA_ALLOC_CONF"), "backend", torch.cuda.memory.get_allocator_backend())
print({k: v for k, v in os.environ.items() if "TORCH" in k or "CUDA" in k})
def stat(): 
    s = torch.cuda.memory_stats(); return s.get("num_device_alloc", -1), s.get("num_device_free", -1), s["reserved_bytes.all.current"] >> 20
B = 1_000_000
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    xs = [torch.empty((7, 3, B), dtype=torch.float64, device="cuda"), torch.empty((2, 3, B), dtype=torch.float64, device="cuda"),
          torch.empty((2, 3, B), dtype=torch.float64, device="cuda"), torch.empty((2, B), dtype=torch.float64, device="cuda"),
          torch.empty((B,), dtype=torch.float64, device="cuda"), torch.empty((3, B), dtype=torch.float64, device="cuda"),
          torch.empty((B,), dtype=torch.int64, device="cuda"), torch.empty((B,), dtype=torch.int64, device="cuda"), torch.empty((B,), dtype=torch.int32, device="cuda")]
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print("alloc 9 tensors", 1e3 * (t1 - t0), "ms", stat())
    del xs
    torch.cuda.synchronize(); print("after del", stat())
# same but touching with a kernel in between
for rep in range(3):
    t0 = time.perf_counter(); x = torch.empty((7, 3, B), dtype=torch.float64, device="cuda"); t1 = time.perf_counter()
    x.zero_(); torch.cuda.synchronize(); t2 = time.perf_counter(); del x
    print("empty", 1e3 * (t1 - t0), "zero_", 1e3 * (t2 - t1), stat())
