import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import subprocess
HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libhm_debug.so")
if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(os.path.join(HERE, "hm_debug.cu")):
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17",
                           "-Xcompiler", "-fPIC", "-shared", "-o", SO, os.path.join(HERE, "hm_debug.cu"), "-lcudart"])
lib = C.CDLL(SO)
lib.hm_debug_mma_issue.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
out = torch.zeros(2, dtype=torch.int64, device="cuda")
for n in (16, 64, 256):
    for off in (0, 1, 3, 4, 8):
        count = 1024
        for _ in range(2):
            lib.hm_debug_mma_issue(n, count, off, out.data_ptr(), None)
            torch.cuda.synchronize()
        o = out.tolist()
        print("N=%3d a_off=%d rows  issue %.1f cyc/MMA   complete %.1f cyc/MMA" % (n, off, o[0] / count, o[1] / count))

lib.hm_debug_mma2_issue.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
for _ in range(2):
    lib.hm_debug_mma2_issue(1024, out.data_ptr(), None)
    torch.cuda.synchronize()
o = out.tolist()
print("cta_group::2 M=256 N=256: issue %.1f cyc/MMA   complete %.1f cyc/MMA" % (o[0] / 1024, o[1] / 1024))

lib.hm_debug_mma_issue_mn.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
for n in (64, 128, 256):
    for lbo in (8192, 128):     # 8192: separate 64-channel boxes; 128: the "next tap = next pixel row" trick of hm_mnrows
        for _ in range(2):
            lib.hm_debug_mma_issue_mn(n, 1024, lbo, out.data_ptr(), None)
            torch.cuda.synchronize()
        o = out.tolist()
        print("MN-major M=128 N=%3d LBO=%4d: issue %.1f cyc/MMA   complete %.1f cyc/MMA" % (n, lbo, o[0] / 1024, o[1] / 1024))
