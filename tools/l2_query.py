import ctypes
rt = ctypes.CDLL("libcudart.so")
def attr(a):
    v = ctypes.c_int(0)
    rt.cudaDeviceGetAttribute(ctypes.byref(v), a, 0)
    return v.value
print("l2", attr(38), "maxPersistingL2", attr(108), "maxAccessPolicyWindow", attr(109))
