import ctypes, os, torch
L = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libmembw3.so"))
B, T, V = 256, 373, 5000
nvt = 10
x = torch.randn(B * T * 5120 + 1024, device="cuda")
out = torch.zeros(4, device="cuda"); ms = ctypes.c_float(0)
print(f"{'layout':>12} {'ld':>5} {'rows':>4} {'cta/sm':>6} {'ms':>8} {'GB/s(useful 5000 cols)':>10}")
for tile_major, ld in ((0, 5000), (0, 5056), (0, 5120), (1, 512)):
    for rows in (4, 8, 16):
        for cps in (4, 8):
            rc = L.membw3_run(ctypes.c_void_p(x.data_ptr()), B, T, ctypes.c_longlong(ld), nvt, tile_major, rows, cps, ctypes.c_void_p(out.data_ptr()), ctypes.byref(ms))
            nbytes = B * T * nvt * 512 * 4
            print(f"{'tile-major' if tile_major else 'row-major':>12} {ld:5d} {rows:4d} {cps:6d} {ms.value:8.3f} {nbytes / ms.value / 1e6:10.1f}  rc={rc}")
