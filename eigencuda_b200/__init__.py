"""eigencuda_b200 -- B200-native replacement for EigenCuda's hot path (CudaPipeline::gemm on
CudaMatrix operands, FP64, column-major).  See DESIGN.md.  Importing this package loads the
in-tree CUDA library; it raises if that library has not been built (no CPU fallback)."""
from ._lib import (EC_DEVICES_AUTO, EC_BACKEND_FP64_DMMA, EC_BACKEND_OZAKI_I8, EC_OP_N, EC_OP_T, EC_STREAM_BASIS,
                   EC_STREAM_LEFT, EC_STREAM_LEFT_T, EC_STREAM_RIGHT, EigenCudaError, LIB_PATH)
from .api import (CudaMatrix, CudaPipeline, DeviceGrid, DistMatrix, Index, dgemm_2d, PinnedBuffer, count_available_gpus, debug_hold_sms, host_link_probe, host_link_probe_buffers, host_pin, nccl_version, pick_stream_devices, plan_chunks, plan_dgemm, plan_panels, plan_shard,
                  result_shape)
from .shard import broadcast_fixed, shard_range

__all__ = ["CudaMatrix", "CudaPipeline", "DeviceGrid", "DistMatrix", "dgemm_2d", "Index", "PinnedBuffer", "host_pin", "host_link_probe", "host_link_probe_buffers", "nccl_version", "pick_stream_devices", "plan_chunks", "plan_dgemm", "plan_panels", "plan_shard", "count_available_gpus",
           "result_shape", "EigenCudaError", "broadcast_fixed", "shard_range", "LIB_PATH",
           "EC_OP_N", "EC_OP_T", "EC_STREAM_RIGHT", "EC_STREAM_LEFT", "EC_STREAM_LEFT_T",
           "EC_STREAM_BASIS", "EC_BACKEND_FP64_DMMA", "EC_BACKEND_OZAKI_I8"]
