"""CPU: stream-order race detection for the frame schedulers - the product's OWN host code (csrc/b2r.cu, unchanged) runs
against a recording stand-in for the CUDA runtime (tools/trace/fake_cudart.cpp: "device" memory is host memory, kernels
do not run, every stream / event / copy / launch is logged) and tools/trace/stream_order.py checks that any two device
operations that touch the same bytes, one of them writing, are ordered by stream order, an event or a host
synchronisation.  Covers the default three-stream pipeline, serial frames and the frame ring (depth 2-4, never run on a
GPU in round 1) on one device with frames in flight, a moving camera, pipelined Present, read-backs, stage timers,
unsorted lists and scene changes between frames.  The detector is shown to have teeth by deleting waits from the log."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "build", "trace")
sys.path.insert(0, os.path.join(ROOT, "tools", "trace"))
import stream_order  # noqa: E402

W, H = 320, 240
MODES = ["default", "serial", "novertex", "ring2", "ring3", "ring4"]


@pytest.fixture(scope="module")
def trace_build():
    if not shutil.which("nvcc"):
        pytest.skip("nvcc not available")
    subprocess.check_call(["bash", os.path.join(ROOT, "tools", "trace", "build.sh")], stdout=subprocess.DEVNULL)
    return {"B2R_LIBRARY": os.path.join(OUT, "libb2r_trace.so"), "B2R_TRACE_PARAMS": os.path.join(OUT, "kernel_params.txt")}


@pytest.fixture(scope="module")
def logs(trace_build, tmp_path_factory):
    d = tmp_path_factory.mktemp("traces")
    out = {}
    for mode in MODES:
        path = str(d / f"{mode}.jsonl")
        res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "trace_scenarios", "frames.py"), mode], capture_output=True, text=True, timeout=300,
                             env=dict(os.environ, B2R_TRACE_FILE=path, **trace_build))
        assert res.returncode == 0 and "scenario done" in res.stdout, res.stdout + res.stderr
        out[mode] = path
    return out


@pytest.mark.parametrize("mode", MODES)
def test_no_unordered_conflicting_accesses(logs, mode):
    t = stream_order.Trace(logs[mode])
    races = t.analyse()
    assert not t.warnings, sorted(set(t.warnings))[:5]                      # every pointer argument was resolved to an allocation
    assert not t.bad_launches() and not t.api_errors()
    assert not t.wiring_errors(), t.wiring_errors()[:4]
    sets = t.buffer_sets(0, "tileCount")[:8]                  # the first frames, all enqueued back to back
    depth = {"default": 2, "serial": 1, "novertex": 1, "ring2": 2, "ring3": 3, "ring4": 4}[mode]
    assert len(set(sets)) == depth and all(sets[i] == sets[i + depth] for i in range(len(sets) - depth)), [hex(x) for x in sets]
    assert not t.leaks(), f"allocations alive after b2r_destroy: {list(t.leaks().values())[:6]}"
    launches = [o for o in t.ops if o["op"] == "launch"]
    assert len(launches) > 300 and all(o["args"] for o in launches)
    streams = {o["stream"] for o in launches}
    assert len(streams) >= (1 if mode in ("serial", "novertex") else 3), streams    # the frames really were spread over streams
    assert not races, "\n".join(t.describe(r) for r in races[:8])


@pytest.mark.parametrize("mode", ["default", "ring3"])
def test_detector_has_teeth(logs, mode, tmp_path):
    lines = open(logs[mode]).readlines()

    def races_without(drop):
        p = tmp_path / "mutated.jsonl"
        p.write_text("".join(l for i, l in enumerate(lines) if i not in drop))
        return len(stream_order.Trace(str(p)).analyse())
    waits = [i for i, l in enumerate(lines) if '"stream_wait"' in l]
    syncs = [i for i, l in enumerate(lines) if '"event_sync"' in l or '"stream_sync"' in l]
    assert len(waits) > 50
    assert races_without(set(waits)) > 100, "frames in flight without any event wait must race"
    assert races_without(set(syncs)) > 10, "read-backs without host synchronisation must race"
    # most single waits are load-bearing (the rest are implied by a host synchronisation or another wait)
    needed = sum(races_without({w}) > 0 for w in waits[::3])
    assert needed >= len(waits[::3]) // 3, (needed, len(waits[::3]))


def test_allocation_time_fills_are_completed_before_use(logs):
    """cudaMemset runs on the legacy default stream, which the context's non-blocking streams are not ordered behind: the
    fills at allocation time must be followed by a synchronisation of that stream (b2r.cu: finishFills)."""
    t = stream_order.Trace(logs["default"])
    ops = t.ops
    fills = [i for i, o in enumerate(ops) if o["op"] == "memset"]
    assert fills
    for i in fills:
        nxt = next(o for o in ops[i + 1:] if o["op"] not in ("memset", "malloc"))
        # (a fill on one of the context's own streams - cudaMemsetAsync: the key buffer of the direct path - is ordered by that
        # stream: it must be synchronised, or followed by more work of the same stream, before anything else)
        assert nxt["op"] in ("stream_sync", "launch", "memcpy") and nxt["stream"] == ops[i]["stream"], (ops[i], nxt)


# ---- the sharded sort-first front end: N contexts of one process stand for the N ranks (tests/trace_scenarios/sharded.py)
SHARDED = [(2, "default"), (2, "serial"), (2, "ring2"), (2, "ring3"), (2, "ring3sliced"), (3, "ring2"), (4, "default"), (4, "ring4sliced"), (8, "default"), (8, "ring3sliced")]


@pytest.fixture(scope="module")
def sharded_logs(trace_build, tmp_path_factory):
    d = tmp_path_factory.mktemp("sharded_traces")
    out = {}
    for world, mode in SHARDED:
        path = str(d / f"{world}_{mode}.jsonl")
        res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "trace_scenarios", "sharded.py"), str(world), mode], capture_output=True, text=True,
                             timeout=300, env=dict(os.environ, B2R_TRACE_FILE=path, B2R_TRACE_DEVICES=str(world), **trace_build))
        assert res.returncode == 0 and "scenario done" in res.stdout, res.stdout + res.stderr
        out[(world, mode)] = path
    return out


@pytest.mark.parametrize("world,mode", SHARDED)
def test_sharded_front_end_has_no_unordered_conflicting_accesses(sharded_logs, world, mode):
    """Lists pushed into the peers' arenas, list sets re-used by later frames, colours stored into rank 0's surface and
    copied to the host by a pipelined Present: every conflicting pair is ordered by the streams, the events and the
    flag protocol (lists ready / colours landed / frame done) - default pipeline and the frame ring with the sliced vertex stage."""
    t = stream_order.Trace(sharded_logs[(world, mode)])
    races = t.analyse()
    assert not t.warnings, sorted(set(t.warnings))[:5]
    assert {o.get("dev", 0) for o in t.ops if o["op"] == "launch"} == set(range(world))
    assert not t.bad_launches() and not t.api_errors()
    assert not t.wiring_errors(world), t.wiring_errors(world)[:4]
    assert not t.leaks(), f"allocations alive after b2r_shard_release + b2r_destroy: {list(t.leaks().values())[:6]}"
    pushes = [o for o in t.ops if o["op"] == "launch" and "bin_scatter" in o["kernel"]]     # the scatter kernel mirrors the lists into the owners' arenas
    assert len(pushes) >= 20 * world
    assert sum("vertex_transform" in o["kernel"] for o in t.ops if o["op"] == "launch") >= 40      # also with the sliced vertex stage (ranks with an empty slice launch none)
    assert not races, "\n".join(t.describe(r) for r in races[:8])


@pytest.mark.parametrize("world,mode", [(2, "default"), (2, "ring3"), (4, "ring4sliced")])
def test_every_element_of_the_flag_protocol_is_load_bearing(sharded_logs, world, mode, tmp_path):
    import json
    lines = open(sharded_logs[(world, mode)]).readlines()

    def barrier(line, base):
        if not ('"launch"' in line and "peer_barrier" in line):
            return False
        b = stream_order.u32(json.loads(line)["args"][3])
        return b == base or (base == 0 and b >= 32)          # "lists ready": words 0..7, or 32 + 8 k for ring slot k

    def races_without(pred):
        p = tmp_path / "mutated.jsonl"
        p.write_text("".join(l for l in lines if not pred(l)))
        return len(stream_order.Trace(str(p)).analyse())
    assert races_without(lambda l: barrier(l, 0)) > 50, "without the 'lists ready' barrier a tile owner reads lists that are still being pushed"
    assert races_without(lambda l: barrier(l, 8)) > 0, "without the 'colours landed' wait rank 0 presents a frame its peers are still writing"
    assert races_without(lambda l: "host_barrier" in l) > 0, "without the host barrier after bind a peer's first stores race with the arena's initial fill"
    if mode.startswith("ring"):
        # the 'frame done' wait is the tail of the frame's clear kernel (its wait-ranks argument): switch it off in the log
        def no_done_wait(line):
            if '"launch"' in line and "frame_clear_kernel" in line:
                o = json.loads(line)
                o["args"][12] = "00000000"
                return json.dumps(o) + "\n"
            return line
        p = tmp_path / "mutated.jsonl"
        p.write_text("".join(no_done_wait(l) for l in lines))
        assert len(stream_order.Trace(str(p)).analyse()) > 0, "without the 'frame done' wait a list set is overwritten while a peer still rasterises from it"


@pytest.mark.parametrize("world,mode,tris", [(1, "default", 0), (1, "ring3", 0), (2, "default", 150), (2, "ring3sliced", 150), (8, "default", 150),
                                             (8, "ring3sliced", 150), (8, "ring2sliced", 1500)])
def test_degenerate_work_sizes(trace_build, tmp_path, world, mode, tris):
    """Empty scene; fewer primitive blocks than ranks (ranks with an empty slice, and no vertex chunk when the vertex stage is
    sliced): every launch has a valid configuration and the frames are still ordered."""
    path = str(tmp_path / "log.jsonl")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "trace_scenarios", "tiny.py"), str(world), mode, str(tris)], capture_output=True, text=True,
                         timeout=300, env=dict(os.environ, B2R_TRACE_FILE=path, B2R_TRACE_DEVICES=str(world), **trace_build))
    assert res.returncode == 0 and "done" in res.stdout, res.stdout + res.stderr
    t = stream_order.Trace(path)
    races = t.analyse()
    assert not t.bad_launches(), t.bad_launches()
    assert not t.warnings, sorted(set(t.warnings))[:5]
    assert not t.wiring_errors(world), t.wiring_errors(world)[:4]
    # (no host barrier marker in this scenario: the arena's initial fill vs the peers' first stores is the host barrier's job)
    races = [r for r in races if not t.nodes[r[0]]["label"].startswith("memset")]
    assert not races, "\n".join(t.describe(r) for r in races[:8])


@pytest.mark.parametrize("argv", [["--steps", "6", "--warmup", "3", "--no-cpu-baseline"],
                                  ["--workload", "views172", "--steps", "8"],
                                  ["--workload", "vehicle1080", "--steps", "5", "--no-cpu-baseline"]], ids=["spheres10m+fill64", "views172", "vehicle1080"])
def test_bench_gpu_arm_end_to_end_on_the_tracing_build(trace_build, tmp_path, argv):
    """bench.py with the REAL ctypes Renderer and C ABI (only torch.cuda is replaced): every call bench.py makes is marshalled
    correctly, one JSON line comes out, and bench's own usage pattern (timed loop, unsorted-list frames, pipelined Present
    loop, stage timers, fragment counting, the fill test's second context) is race-free."""
    import json
    path = str(tmp_path / "log.jsonl")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "trace_scenarios", "bench_dry.py"), *argv], capture_output=True, text=True, timeout=600,
                         env=dict(os.environ, B2R_TRACE_FILE=path, **trace_build))
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-3000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["value"] > 0 and d["e2e"]["value"] > 0 and d["gpu_launches"] > 0
    t = stream_order.Trace(path)
    races = t.analyse()
    assert not t.warnings and not t.bad_launches() and not t.api_errors() and not t.leaks()
    assert not races, "\n".join(t.describe(r) for r in races[:8])


def test_api_misuse_returns_status_codes(trace_build, tmp_path):
    """69 wrong calls (order, ids, null pointers, sizes, options at the wrong time): each returns a negative status with a
    message, none crashes, and the context renders a normal frame afterwards."""
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "trace_scenarios", "api_misuse.py")], capture_output=True, text=True, timeout=300,
                         env=dict(os.environ, B2R_TRACE_FILE=str(tmp_path / "log.jsonl"), **trace_build))
    assert res.returncode == 0 and "checks passed" in res.stdout, res.stdout[-3000:] + res.stderr[-2000:]
    assert "FAIL" not in res.stdout


@pytest.mark.skipif(os.environ.get("B2R_SLOW_TESTS") != "1", reason="builds an AddressSanitizer variant of the tracing library (~1 min; set B2R_SLOW_TESTS=1)")
def test_host_code_is_clean_under_address_sanitizer(tmp_path):
    """The product's host code (schedulers, uploads, the veneer's OBJ parser / PNG decoder, error paths) under ASan + UBSan."""
    asan = subprocess.run(["g++", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(asan) or not shutil.which("nvcc"):
        pytest.skip("libasan / nvcc not available")
    subprocess.check_call(["bash", os.path.join(ROOT, "tools", "trace", "build.sh"), "--asan"], stdout=subprocess.DEVNULL)
    out = os.path.join(ROOT, "build", "trace_asan")
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0", B2R_LIBRARY=os.path.join(out, "libb2r_trace.so"),
               B2R_TRACE_PARAMS=os.path.join(out, "kernel_params.txt"))
    runs = [("frames.py", ["default"], 1), ("frames.py", ["ring3"], 1), ("sharded.py", ["2", "default"], 2), ("sharded.py", ["4", "ring3sliced"], 4),
            ("tiny.py", ["8", "ring3sliced", "150"], 8), ("api_misuse.py", [], 1)]
    for script, args, devices in runs:
        res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "trace_scenarios", script), *args], capture_output=True, text=True, timeout=600,
                             env=dict(env, B2R_TRACE_FILE=str(tmp_path / "log.jsonl"), B2R_TRACE_DEVICES=str(devices)))
        assert res.returncode == 0 and "AddressSanitizer" not in res.stderr and "runtime error" not in res.stderr, (script, args, res.stdout[-800:], res.stderr[-3000:])
    res = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_cpu_host.py"), "-q", "-x", "-p", "no:cacheprovider",
                          "-k", "objparser or png or camera or mesh_world or null_context or bad_resolution"], capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert res.returncode == 0 and "AddressSanitizer" not in res.stderr, res.stdout[-2000:] + res.stderr[-3000:]


@pytest.mark.skipif(os.environ.get("B2R_SLOW_TESTS") != "1", reason="builds a ThreadSanitizer variant of the tracing library (~1 min; set B2R_SLOW_TESTS=1)")
def test_independent_contexts_on_concurrent_threads_under_thread_sanitizer():
    """"Thread-compatible per handle": four host threads drive four contexts at the same time (uploads, frames in flight, ring
    option, read-backs, destroy) while a fifth exercises b2r_create's per-thread error text; TSan sees no shared mutable state."""
    if not shutil.which("nvcc"):
        pytest.skip("nvcc not available")
    subprocess.check_call(["bash", os.path.join(ROOT, "tools", "trace", "build.sh"), "--tsan"], stdout=subprocess.DEVNULL)
    out = os.path.join(ROOT, "build", "trace_tsan")
    res = subprocess.run([os.path.join(out, "threads")], capture_output=True, text=True, timeout=600, env=dict(os.environ, B2R_TRACE_PARAMS=os.path.join(out, "kernel_params.txt")))
    assert res.returncode == 0 and "0 failures" in res.stdout and "ThreadSanitizer" not in res.stderr, res.stdout[-500:] + res.stderr[-3000:]


def test_wiring_check_has_teeth(logs, sharded_logs, tmp_path):
    """A raster kernel given the lists of the other buffer set, and a tile owner given a peer's heads from the wrong list
    set, are ordered accesses (no race) - the pointer-level dataflow check is what catches them."""
    import json

    def mutate(path, pick, patch):
        lines = open(path).readlines()
        idx = [i for i, l in enumerate(lines) if '"launch"' in l and "raster_shade" in l and pick(l)]
        a, b = json.loads(lines[idx[4]]), json.loads(lines[idx[5]])
        raw = bytearray(bytes.fromhex(a["args"][0]))
        patch(raw, bytes.fromhex(b["args"][0]))
        a["args"][0] = raw.hex()
        lines[idx[4]] = json.dumps(a) + "\n"
        p = tmp_path / "mutated.jsonl"
        p.write_text("".join(lines))
        return stream_order.Trace(str(p))

    def other_lists(raw, nxt):
        raw[16:24] = nxt[16:24]                  # RasterArgs.binIds of the next frame (the other set)
    t = mutate(logs["default"], lambda l: True, other_lists)
    assert any("binIds" in e for e in t.wiring_errors())

    def other_inbox(raw, nxt):
        raw[15 * 8:16 * 8] = nxt[15 * 8:16 * 8]  # RasterArgs.shard.tileCount[0] of the next frame's list set
    t = mutate(sharded_logs[(2, "ring3")], lambda l: '"dev":1' in l, other_inbox)
    assert any("reads the heads of rank 0" in e for e in t.wiring_errors(2))


@pytest.mark.parametrize("graph,cam,total", [(3, (0, 1, 5), 292), (2, (0, 5, 65), 4878), (4, (0, 5, 65), 11638)], ids=["bunny", "tuktuk", "vehicle"])
def test_reference_scene_code_drives_the_renderer_on_the_tracing_build(trace_build, tmp_path, graph, cam, total):
    """oracle/_ref/dropin_main (the reference's own CreateScenes + the veneer, see tests/test_gpu_veneer.py) through LoadScene and
    RenderAuto on the recording runtime: the call sequence of the reference's main() completes, uploads the scene graph's
    triangles and is race-free.  (The frames themselves are compared on a GPU.)"""
    exe = os.path.join(ROOT, "oracle", "_ref", "dropin_main")
    res_dir = os.path.join(ROOT, "oracle", "_ref", "resources")
    if not (os.path.exists(exe) and os.path.exists(os.path.join(res_dir, "vehicle.obj"))):
        pytest.skip("oracle/_ref/dropin_main not built (needs /root/reference at build time)")
    os.symlink(res_dir, tmp_path / "Resources")
    log, out = str(tmp_path / "log.jsonl"), str(tmp_path / "o.raw")
    preload = trace_build["B2R_LIBRARY"] + ":" + os.path.join(OUT, "libfake_cudart.so")       # the binary is linked against the real libb2r.so
    res = subprocess.run([exe, str(tmp_path), "render", str(graph), "640", "480", *[str(c) for c in cam], "0", "1", out], capture_output=True, text=True, timeout=300,
                         env=dict(os.environ, LD_PRELOAD=preload, B2R_TRACE_FILE=log, **trace_build))
    assert res.returncode == 0 and f"triangles {total} visible" in res.stdout, res.stdout + res.stderr
    assert os.path.getsize(out) == 640 * 480 * 4
    t = stream_order.Trace(log)
    assert not t.analyse() and not t.bad_launches() and not t.wiring_errors() and not t.leaks()
