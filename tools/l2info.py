import torch, ctypes
p = torch.cuda.get_device_properties(0)
print("L2", p.L2_cache_size >> 20, "MB")
rt = ctypes.CDLL("libcudart.so.12") if False else None
from cuda import cudart
for name in ("cudaDevAttrMaxPersistingL2CacheSize", "cudaDevAttrMaxAccessPolicyWindowSize", "cudaDevAttrL2CacheSize"):
    err, v = cudart.cudaDeviceGetAttribute(getattr(cudart.cudaDeviceAttr, name), 0)
    print(name, v >> 20, "MB")
