"""CPU: the product's KERNELS, executed.  tools/emu builds the product's sources with the host compiler on top of a small CUDA
execution model (tools/emu/emu_cuda.h: every CUDA thread of a block is a fiber; __syncthreads and the *_sync warp
collectives are rendezvous points; atomics, shared memory, the built-in variables and the intrinsics the kernels use)
and the host-memory runtime of tools/trace.  The kernel sources are the product's, byte for byte (csrc/*.cuh); only the
launch syntax of b2r.cu, its inline PTX and one MUFU instruction are rewritten (tools/emu/preprocess.py).  It is test
infrastructure - the product library has no CPU path - and it is slow (seconds per small frame), but it lets the whole
pipeline be compared with the oracle without a GPU, in particular the scheduler modes that no GPU was available for:

* parity: transformed vertices, visibility, per-tile lists (sets AND order), winners, integer depth, interpolated z and
  colour of the emulated product equal the oracle's bit for bit on the golden frame cases;
* schedulers: seven frames with a moving camera enqueued back to back with a pipelined Present each - default three-stream
  pipeline, serial frames, unsorted lists canonicalised on read, frame ring of depth 2 / 3 / 4 - every presented frame
  equals the oracle's.  (A launch executes immediately, block by block: one valid serialisation of the stream order on one
  device.  That frames in flight are ORDERED is tests/test_cpu_trace.py's job; that they compute the right thing is this
  file's.)  Streams are real queues here (tools/emu/emu_runtime.cpp): work is executed when the host synchronises, newest
  stream first, so a later frame's front end runs before an earlier frame's raster wherever the product's events allow
  it - with the event waits switched off (B2R_EMU_IGNORE_WAITS) every presented frame comes out wrong;
* sharded sort-first front end: N contexts of one process as the N ranks (the flag barriers work because a kernel that
  sleeps in its spin loop hands the processor to the other streams); seven frames in flight, composed on rank 0 - default
  pipeline, frame ring, frame ring + sliced vertex stage (the raster kernel transforming its vertices itself): every
  composed frame equals the oracle's and the ranks' lists concatenate to the oracle's lists."""
import json
import os
import shutil
import subprocess
import sys

import pytest

import golden_cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "build", "emu", "libb2r_emu.so")


@pytest.fixture(scope="module")
def emu_env():
    if not shutil.which("g++") or not os.path.exists("/usr/local/cuda/include/cuda_runtime.h"):
        pytest.skip("host compiler / CUDA headers not available")
    subprocess.check_call(["bash", os.path.join(ROOT, "tools", "emu", "build.sh")], stdout=subprocess.DEVNULL)
    return dict(os.environ, B2R_LIBRARY=LIB, B2R_EMU_TESTS="1")      # tests/conftest.py refuses -m gpu on the emulation otherwise


def run(env, *args):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emu_scenarios.py"), *args], capture_output=True, text=True, timeout=900, env=env)
    assert res.returncode == 0, res.stdout[-1000:] + res.stderr[-3000:]
    return json.loads(res.stdout.strip().splitlines()[-1])


def all_true(d):
    return all(v is True or (k == "colour_max_diff" and v == 0) or (k == "kernels" and v >= 9) for k, v in d.items())


@pytest.mark.parametrize("case", sorted(golden_cases.frame_cases()))
def test_emulated_product_equals_the_oracle(emu_env, case):
    res = run(emu_env, "parity", case)
    assert all_true(res), res


@pytest.mark.parametrize("case", ["soup_colour", "quads_linear", "hostile"])
def test_emulated_product_equals_the_oracle_with_the_reference_fill_rule(emu_env, case):
    res = run(emu_env, "parity", case, "0")
    assert all_true(res), res


SLOW = os.environ.get("B2R_SLOW_TESTS") == "1"          # the full matrix takes ~10 min of CPU; the default subset ~4 min


@pytest.mark.parametrize("mode", ["default", "sorted", "ring3"] + (["unsorted", "serial", "ring2", "ring4"] if SLOW else []))
def test_scheduler_modes_compute_the_right_frames(emu_env, mode):
    res = run(emu_env, "schedulers", mode)
    assert all(res["frames"]) and len(res["frames"]) == 7, res["frames"]
    assert all_true(res["last"]), res["last"]
    assert all(all_true(s) for s in res["stepwise"]), res["stepwise"]


def test_direct_path_equals_the_binned_path(emu_env):
    """B2R_OPT_DIRECT_MAX_AREA (small triangles resolved by the setup kernel with atomicMin on the key buffer) against the
    all-binned path: surfaces, statistics and complete bins identical for several bin grids, thresholds, fill rules,
    sort-first ownership and with frames in flight (tests/emu_scenarios.py direct_path_checks)."""
    # (B2R_SETUP_GRID: two CTAs for the setup kernel, so that each walks several 256-primitive chunks through its copy pipeline)
    res = run(dict(emu_env, B2R_SETUP_GRID="2"), "direct", "full" if SLOW else "quick")
    assert res and all(res.values()), {k: v for k, v in res.items() if not v}
    assert "inflight/ring3" in res and "owners3" in res


@pytest.mark.parametrize("mode", ["ring3"] + (["default"] if SLOW else []))
def test_the_emulation_is_sensitive_to_the_event_waits(emu_env, mode):
    res = run(dict(emu_env, B2R_EMU_IGNORE_WAITS="1"), "schedulers", mode)
    assert not any(res["frames"]), "frames in flight without their event waits must not come out right"


# (ranks, mode, meshes): 7 meshes - the raster kernel's on-the-fly vertex transform keeps the matrices of the first 4 meshes in
# shared memory and loads the others from global memory
# 150 triangles: one 256-primitive block, i.e. all ranks but one own an EMPTY primitive slice
SHARDED = [(2, "ring3sliced", 7, 700), (4, "ring4sliced", 2, 150)]
if SLOW:
    SHARDED += [(2, "default", 2, 700), (2, "serial", 2, 700), (2, "ring3", 2, 700), (3, "ring2", 2, 700), (8, "default", 2, 700), (8, "ring3sliced", 2, 700),
                (8, "default", 1, 150), (8, "ring3sliced", 1, 150), (4, "ring2sliced", 3, 300), (4, "ring4sliced", 2, 700)]


@pytest.mark.parametrize("world,mode,meshes,tris", SHARDED)
def test_sharded_front_end_composes_the_right_frames(emu_env, world, mode, meshes, tris):
    res = run(dict(emu_env, B2R_TRACE_DEVICES=str(world)), "sharded", str(world), mode, str(meshes), str(tris))
    assert all(res["frames"]) and len(res["frames"]) == 7, res
    assert res["last_framebuffer"] and res["lists"] and res["visible"], res
    assert res["distributed_present"] and res["band_present"] and res["composed_on_rank0_again"], res


@pytest.mark.parametrize("world,mode,meshes,tris", [(4, "ring4sliced", 2, 400)] + ([(2, "default", 2, 500), (8, "ring3", 3, 500), (3, "serial", 1, 400)] if SLOW else []))
def test_sharded_front_end_with_small_triangles_through_the_key_buffers(emu_env, world, mode, meshes, tris):
    """A scene of mostly few-pixel triangles: the ranks resolve them into their own key buffers (direct path) and push the
    touched 16x16 key blocks to the tile owners; with the sliced vertex stage the raster kernel transforms the winners'
    vertices itself.  Every composed frame equals the oracle's; the statistics add up across the ranks."""
    res = run(dict(emu_env, B2R_TRACE_DEVICES=str(world), B2R_EMU_SCENE="tiny"), "sharded", str(world), mode, str(meshes), str(tris))
    assert all(res["frames"]) and len(res["frames"]) == 7, res
    assert res["last_framebuffer"] and res["last_framebuffer_binned"] and res["lists"] and res["visible"] and res["visible_direct"] and res["pairs_direct"], res
    assert res["distributed_present"] and res["band_present"] and res["composed_on_rank0_again"], res


@pytest.mark.skipif(not SLOW, reason="other stream visiting orders of the emulated executor (set B2R_SLOW_TESTS=1)")
@pytest.mark.parametrize("order", ["asc", "random:3", "random:11"])
@pytest.mark.parametrize("mode", ["default", "ring3"])
def test_scheduler_modes_under_other_execution_orders(emu_env, mode, order):
    res = run(dict(emu_env, B2R_EMU_ORDER=order), "schedulers", mode)
    assert all(res["frames"]) and all_true(res["last"]) and all(all_true(s) for s in res["stepwise"]), res


def test_smoke_on_the_emulation(emu_env):
    """__graft_entry__.smoke() - the round-end gate's first call on the GPU box - end to end on the emulation build."""
    res = subprocess.run([sys.executable, "-c", "import __graft_entry__ as g; g.smoke()"], capture_output=True, text=True, timeout=900, env=emu_env, cwd=ROOT)
    assert res.returncode == 0 and "smoke ok [CPU EMULATION BUILD" in res.stdout, res.stdout[-1500:] + res.stderr[-3000:]


def test_cpp_veneer_and_drop_in_gpu_tests_on_the_emulation(emu_env):
    """tests/test_gpu_veneer.py (-m gpu) as it is, with the emulation build standing in for the GPU: the C++ veneer example and
    oracle/_ref/dropin_main (the reference's own CreateScenes on the veneer; the binaries are linked against the real
    libb2r.so, hence the preload) render bunny, TukTuk and vehicle exactly like the C-ABI path."""
    preload = LIB + ":" + os.path.join(os.path.dirname(LIB), "libemu_runtime.so")
    res = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_veneer.py"), "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider"],
                         capture_output=True, text=True, timeout=2400, env=dict(emu_env, LD_PRELOAD=preload), cwd=ROOT)
    assert res.returncode == 0 and " passed" in res.stdout and "failed" not in res.stdout, res.stdout[-3000:] + res.stderr[-2000:]
    assert "skipped" not in res.stdout or not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "dropin_main")), res.stdout[-500:]


def test_gpu_unit_tests_on_the_emulation(emu_env):
    """tests/test_gpu_units.py: every tier of the list sorter (warp run ordering, CTA run ordering / radix, the long radix sort
    in dynamic shared memory) on lists of up to tens of thousands of entries - as it is, on the emulation build; and the
    hoisted division against IEEE division on 2^22 operand pairs (here the refinement starts from the correctly rounded
    reciprocal instead of MUFU.RCP; the 2^32-pair check against div.rn on hardware stays a -m gpu test)."""
    res = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_units.py"), "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider", "-k", "list_sort"],
                         capture_output=True, text=True, timeout=2400, env=emu_env, cwd=ROOT)
    assert res.returncode == 0 and "3 passed" in res.stdout, res.stdout[-3000:] + res.stderr[-2000:]
    code = ("import ctypes as C\nfrom custom_gpgpu_rasterizer_b200._lib import lib\nbad = C.c_uint64(1)\n"
            "assert lib.b2r_selftest_division(0, 1 << 22, 7, C.byref(bad)) == 0 and bad.value == 0, bad.value\nprint('division ok')")
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900, env=emu_env, cwd=ROOT)
    assert res.returncode == 0 and "division ok" in res.stdout, res.stdout + res.stderr[-2000:]
    # the sign-only back-face test against the reference's normalise + dot on 2^23 adversarial decisions per seed (here with
    # glibc's correctly rounded sqrtf / division standing in for sqrt.rn / rcp.rn; 2^33 per seed on hardware is a -m gpu test)
    code = ("import ctypes as C\nfrom custom_gpgpu_rasterizer_b200._lib import lib\n"
            "for seed in (1, 2, 3):\n"
            "    counts = (C.c_uint64 * 2)(7, 0)\n"
            "    assert lib.b2r_selftest_cull(0, 1 << 22, seed, counts) == 0 and counts[0] == 0 and counts[1] > (1 << 20), list(counts)\n"
            "print('cull ok')")
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900, env=emu_env, cwd=ROOT)
    assert res.returncode == 0 and "cull ok" in res.stdout, res.stdout + res.stderr[-2000:]


def test_raster_grid_hint_and_tail_on_the_emulation(emu_env):
    """tests/test_gpu_parity.py::test_raster_grid_follows_the_frame_and_the_tail_catches_the_rest as it is: a camera jump
    makes a frame's work items outgrow the grid the earlier frames suggested, and the tail kernel has to render the rest."""
    res = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider", "-k", "raster_grid_follows" + ("" if SLOW else " and 3")],
                         capture_output=True, text=True, timeout=2400, env=emu_env, cwd=ROOT)
    assert res.returncode == 0 and ("2 passed" if SLOW else "1 passed") in res.stdout, res.stdout[-3000:] + res.stderr[-2000:]


@pytest.mark.skipif(os.environ.get("B2R_EMU_FULL") != "1", reason="~20 min of CPU: the whole -m gpu parity file on the emulation build (set B2R_EMU_FULL=1)")
def test_whole_gpu_parity_file_on_the_emulation(emu_env):
    """tests/test_gpu_parity.py as it is (64 tests: every scene in both fill rules incl. the bundled vehicle at 1080p, the 172-instance
    views with 2 M triangles and the 4K slivers, the hostile-input fuzz, pipelined frames, pipelined Present, list-capacity regrowth)
    with the emulation build standing in for the GPU.  Last full run: 64 passed in 17 min 26 s (round 1, session 4)."""
    res = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider"],
                         capture_output=True, text=True, timeout=3600, env=emu_env, cwd=ROOT)
    assert res.returncode == 0 and " passed" in res.stdout and "failed" not in res.stdout, res.stdout[-3000:] + res.stderr[-2000:]


def test_bench_refuses_the_emulation_build(emu_env):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "spheres_small", "--steps", "1"], capture_output=True, text=True, timeout=300, env=emu_env)
    assert res.returncode != 0 and res.stdout.strip() == "" and "CPU emulation build" in res.stderr


@pytest.mark.parametrize("world", [3])
def test_replicated_sort_first_composes_the_right_frame(emu_env, world):
    """Tile ownership, the packed tile-major output and b2r_unpack_tiles at a ragged resolution (partial tiles): the frame rank 0
    composes from the ranks' packed buffers is the oracle's, and every (triangle, tile) pair was binned by exactly one owner."""
    res = run(dict(emu_env, B2R_TRACE_DEVICES=str(world)), "replicated", str(world))
    assert res == {"framebuffer": True, "pairs": True}, res


@pytest.mark.parametrize("mode", ["default", "ring3"])
def test_scene_changes_between_frames_in_flight(emu_env, mode):
    """b2r_mesh_set_world and a mesh upload between frames that were enqueued without synchronisation: every presented frame shows
    the scene as it was when the frame was issued (per-mesh constants, triangle references and buffers are re-made in order)."""
    res = run(emu_env, "dynamic", mode)
    assert all(res["frames"]) and len(res["frames"]) == 8, res


@pytest.mark.skipif(not SLOW, reason="AddressSanitizer build of the emulation (~3 min; set B2R_SLOW_TESTS=1)")
def test_kernels_are_memory_safe_under_address_sanitizer(tmp_path):
    """The emulated kernels under ASan: 'device' buffers are heap blocks and __shared__ arrays are statics, so an out-of-bounds
    access of any kernel is reported - the role compute-sanitizer's memcheck plays on a GPU.  Hostile inputs, the frame ring,
    the sharded + sliced front end with 7 meshes (stores into the peers' arenas) and the CTA-level sorters."""
    asan = subprocess.run(["g++", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(asan):
        pytest.skip("libasan not available")
    subprocess.check_call(["bash", os.path.join(ROOT, "tools", "emu", "build.sh"), "--asan"], stdout=subprocess.DEVNULL)
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0",
               B2R_LIBRARY=os.path.join(ROOT, "build", "emu_asan", "libb2r_emu.so"), B2R_EMU_TESTS="1")
    for args, devices in ((["parity", "hostile_textured"], 1), (["parity", "soup_textured_linear"], 1), (["schedulers", "ring3"], 1), (["sharded", "2", "ring3sliced", "7"], 2)):
        res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emu_scenarios.py"), *args], capture_output=True, text=True, timeout=1800,
                             env=dict(env, B2R_TRACE_DEVICES=str(devices)))
        assert res.returncode == 0 and "ERROR: AddressSanitizer" not in res.stderr, (args, res.stdout[-500:], res.stderr[-3000:])
    res = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_units.py") + "::test_list_sort_tiers_give_canonical_lists[res1-grid1-tess1]",
                          "-m", "gpu", "-q", "-p", "no:cacheprovider"], capture_output=True, text=True, timeout=1800, env=env, cwd=ROOT)
    assert res.returncode == 0 and "ERROR: AddressSanitizer" not in res.stderr, res.stdout[-1500:] + res.stderr[-3000:]


@pytest.mark.skipif(not SLOW, reason="set B2R_SLOW_TESTS=1")
@pytest.mark.parametrize("world,mode", [(2, "ring3sliced"), (4, "default")])
def test_sharded_front_end_with_hostile_inputs(emu_env, world, mode):
    """Degenerate / non-finite / camera-plane-crossing triangles through the sharded front end, also with the raster kernel
    transforming its vertices itself: the composed frames and the concatenated lists are still the oracle's."""
    res = run(dict(emu_env, B2R_TRACE_DEVICES=str(world), B2R_EMU_SCENE="hostile"), "sharded", str(world), mode)
    assert all(res["frames"]) and res["last_framebuffer"] and res["lists"] and res["visible"], res
