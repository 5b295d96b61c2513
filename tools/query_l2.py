import ctypes, torch
torch.cuda.init()
rt = ctypes.CDLL("libcudart.so")
def attr(a):
    v = ctypes.c_int(0); rc = rt.cudaDeviceGetAttribute(ctypes.byref(v), a, 0); return rc, v.value
# cudaDevAttrL2CacheSize=38, MaxPersistingL2CacheSize=108, MaxAccessPolicyWindowSize=109, cudaDevAttrMaxSharedMemoryPerBlockOptin=97
for name, a in (("L2CacheSize",38),("MaxPersistingL2CacheSize",108),("MaxAccessPolicyWindowSize",109),("SmemOptin",97),("SmemPerSM",81)):
    print(name, attr(a))
