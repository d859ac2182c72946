import ctypes, os, torch
L = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libmembw2.so"))
B, W, T, V = 256, 10, 373, 5000
buf = torch.empty(T * 2 * B * W * 5376, dtype=torch.float32, device="cuda")
ms = ctypes.c_float(0)
print(f"{'ld':>5} {'NT':>4} {'NV':>3} {'HW':>3} {'ord':>3} {'cs':>2} {'piece':>7} {'rows':>4} {'ms':>8} {'GB/s':>8}")
cases = []
import sys
if len(sys.argv) > 1:
    for ld in (5000, 5004, 5008, 5016, 5024, 5032, 5056, 5088, 5120, 5248, 5376):
        cases.append((ld, 128, 1, 5, 0, 1))
        cases.append((ld, 128, 1, 5, 0, 0))
for ld in (() if len(sys.argv) > 1 else (5000, 5120)):
    for (NT, NV, HW) in [(128, 1, 5), (256, 1, 5), (128, 2, 5), (256, 2, 5), (128, 5, 2), (256, 5, 1), (256, 5, 2), (512, 3, 1), (128, 10, 1), (128, 1, 10), (128, 1, 1), (128, 1, 2)]:
        for order in (0, 1, 2):
            cases.append((ld, NT, NV, HW, order, 1))
for (ld, NT, NV, HW, order, cs) in cases:
    rc = L.membw2_run(ctypes.c_void_p(buf.data_ptr()), B, W, T, V, ld, NT, NV, HW, order, cs, ctypes.byref(ms))
    nbytes = T * 2 * B * W * V * 4
    print(f"{ld:5d} {NT:4d} {NV:3d} {HW:3d} {order:3d} {cs:2d} {NT*NV*16:7d} {HW*2:4d} {ms.value:8.3f} {nbytes / ms.value / 1e6:8.1f}  rc={rc}")
