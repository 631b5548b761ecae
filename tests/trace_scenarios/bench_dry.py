"""bench.py's GPU arm end to end on a machine without a GPU: the REAL ctypes Renderer on the tracing build of the library
(B2R_LIBRARY = build/trace/libb2r_trace.so, recording CUDA runtime), only torch.cuda replaced by stand-ins.  Exercises
the argument marshalling of every C-ABI call bench.py makes (render, present_async, stage timers, counters, options) and
leaves a log for tools/trace/stream_order.py.     python bench_dry.py [bench.py arguments]"""
import contextlib
import importlib.util
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
# the tracing library first: its CUDA-runtime symbols are bound when it is loaded, and torch brings the real libcudart
# into the global symbol scope
from custom_gpgpu_rasterizer_b200 import _lib  # noqa: E402,F401
import torch  # noqa: E402

cpu = torch.device("cpu")


class _Event:
    def __init__(self, enable_timing=False):
        pass

    def record(self, stream=None):
        pass

    def elapsed_time(self, other):
        return 9.4


real_empty = torch.empty
torch.device = lambda *a, **k: cpu
torch.cuda.set_device = lambda i: None
torch.cuda.synchronize = lambda *a: None
torch.cuda.Stream = lambda device=None: types.SimpleNamespace(cuda_stream=0)      # 0 = keep the context's own stream
torch.cuda.Event = _Event
torch.cuda.stream = lambda s: contextlib.nullcontext()
torch.empty = lambda *a, pin_memory=False, **k: real_empty(*a, **k)

spec = importlib.util.spec_from_file_location("b2r_bench", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)
sys.argv = ["bench.py"] + sys.argv[1:]
bench.main()
