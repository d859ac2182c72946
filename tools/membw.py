import ctypes, os, torch
L = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libmembw.so"))
B, W, T, V = 256, 10, 373, 5000
n = T * 2 * B * W * V
buf = torch.empty(n, dtype=torch.float32, device="cuda")
ms = (ctypes.c_float * 16)()
rc = L.membw_run(ctypes.c_void_p(buf.data_ptr()), ctypes.c_size_t(n * 4), B, W, T, V, ms)
names = ["st.global grid-stride 256thr", "st.global.cs", "st.global.wt", "st.global.cg", "st.global 1024thr", "cudaMemset", "scorer pattern st", "scorer pattern st.cs"]
print("rc", rc, "bytes", n * 4 / 1e9, "GB")
for nm, m in zip(names, ms):
    print(f"{nm:32s} {m:8.3f} ms  {n * 4 / m / 1e6:8.1f} GB/s")
# torch copy for reference (read+write)
a = torch.empty(n // 2, dtype=torch.float32, device="cuda"); b = torch.empty_like(a)
for _ in range(2): b.copy_(a)
torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); [b.copy_(a) for _ in range(3)]; e1.record(); torch.cuda.synchronize()
m = e0.elapsed_time(e1) / 3
print(f"torch copy (r+w bytes)           {m:8.3f} ms  {n * 4 / m / 1e6:8.1f} GB/s")
