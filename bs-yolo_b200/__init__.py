"""bs_yolo_b200 -- B200-native (sm_100a) detection inference path for the feiyeha/BS-YOLO Ultralytics fork.

Only the hot path lives here: spec/IR of the detect graphs, the planner that lowers them to fused
NHWC bf16 kernels, the ctypes binding to the C-ABI library (``libb2y.so``, built from ``csrc/``),
the host-side mirror of the reference's predictor/ops interface, and batch sharding across GPUs.
"""
__version__ = "0.1.0"
