"""flake16_framework_b200: the `scores` hot path of flake16-framework on B200 GPUs."""
import os

# The grid engine keeps ~30 CUDA streams busy.  With the driver's default of 8 hardware work
# queues, unrelated streams share a queue and a kernel that waits for resources (a DecisionTree
# CTA needs a whole SM) stalls every stream behind it: measured 6.3 s -> 4.8 s per 54-config
# slice with 32 queues.  The variable is read when the CUDA context is created, so it has to be
# set before the first CUDA call of the process (importing this package first is enough).
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

DEFAULT_WORKERS = 4     # host threads driving (dataset, fold) units; 4 x (4 + 2 x lanes) = 32 streams
DEFAULT_LANES = 2       # side streams per forest model and worker
