"""pinned host -> device copy bandwidth of the box (what bounds LGBM_DatasetPushRows ingestion)."""
import json, sys, time
sys.path.insert(0, ".")
from mmlspark_b200 import capi
nb = 1 << 30
h = capi.PinnedBuffer(nb); d = capi.DeviceBuffer(nb)
capi.memcpy(d.ptr, h.ptr, nb)
t = time.perf_counter()
for _ in range(8):
    capi.memcpy(d.ptr, h.ptr, nb)
dt = time.perf_counter() - t
print(json.dumps({"h2d_pinned_GBps": 8 * nb / dt / 1e9}))
