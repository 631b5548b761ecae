#!/usr/bin/env python
"""Stream-order race detector for the frame schedulers - needs no GPU.

Input: the JSON-lines log the recording CUDA runtime (tools/trace/fake_cudart.cpp) writes while the product's own host
code (build/trace/libb2r_trace.so) runs.  Every device operation (kernel launch, copy, memset) becomes a node with the
memory ranges it reads / writes; edges are what CUDA and the product's flag protocol guarantee:

* order within a stream (the legacy default stream of a device is ordered in itself only: the product's streams are
  non-blocking);
* event record -> stream wait;
* host synchronisation of the issuing context (cudaStreamSynchronize / cudaEventSynchronize / cudaFree / blocking
  copies): what that host thread has waited for precedes everything it issues afterwards.  Several contexts of one
  process stand for the ranks of a sharded group (the log carries the issuing device); a rank's host synchronisation
  orders nothing for the other ranks;
* a host-side barrier between the ranks (log marker written by the scenario where a real group calls
  torch.distributed.barrier);
* the cross-GPU flag protocol of the sharded front end, from the scalar arguments of the kernels that implement it:
  peer_barrier_kernel(signal = 1, wait = 1, slot base b, epoch e) of all ranks form a barrier - everything before the
  kernel on ANY rank precedes everything after it on EVERY rank; publish_info_kernel signals "colours landed" (to rank
  0, flag words 8..) and, with the frame ring, "frame done" (to all ranks, words 16..) with its epoch;
  peer_barrier_kernel(signal = 0, wait = 1, b, e) waits for those signals of all ranks.

Two operations CONFLICT when they touch overlapping bytes, at least one of them writes, and neither is ordered before
the other: a data race between frames in flight (or between ranks), whatever the timing on a real GPU happens to be.

What a kernel touches is derived from its launch arguments: pointer parameters from the demangled signature
(`const T*` = read, `T*` = write), the by-value structs RasterArgs / ShardInbox / ShardRegions from their layout
(csrc/raster_kernel.cuh, csrc/device_types.cuh).  A pointer stands for the bytes from it to the next address that is
ever passed to a kernel as a pointer into the same allocation (or the end of the allocation): sub-buffers carved out of
one allocation - the two colour surfaces, the list sets / inboxes / regions of the shared arena - are told apart,
because the base of every sub-buffer is some kernel's argument.  (A pointer into the MIDDLE of a logical buffer, e.g. a
slice of the vertex chunks, narrows the range of the accesses through the buffer's base: possible false negatives, no
false positives.)  The flag words themselves are accessed concurrently by design (volatile stores / polling) and are
left out.  Programmatic dependent launch does not relax stream order here: every kernel of the product begins with
griddepcontrol.wait (pdlPrologue), i.e. waits for the completion and memory flush of its predecessor.

    python tools/trace/stream_order.py log.jsonl [-v]
exit status 1 when a race is found.  tests/test_cpu_trace.py drives it."""
from __future__ import annotations

import argparse
import bisect
import json
import subprocess
import sys
from collections import defaultdict

K_MAX_RANKS = 8
RASTER_ARGS = [("tileCount", "r"), ("tileOffset", "r"), ("binIds", "w"), ("triRefs", "r"), ("pos4", "r"), ("posIn", "r"), ("ivertex", "r"), ("colour", "r"),
               ("meshes", "r"), ("textures", "r"), ("frameInfo", "w"), ("framebuffer", "w"), ("depth", "w"), ("zSurface", "w"), ("winnerSurface", "w")]
RASTER_ARGS += [(f"shard.tileCount[{i}]", "r") for i in range(K_MAX_RANKS)] + [(f"shard.tileOffset[{i}]", "r") for i in range(K_MAX_RANKS)] \
    + [(f"shard.binIds[{i}]", "w") for i in range(K_MAX_RANKS)] + [(f"shard.keyInbox[{i}]", "r") for i in range(K_MAX_RANKS)] \
    + [(f"shard.dirtyInbox[{i}]", "w") for i in range(K_MAX_RANKS)] + [("tileOrder", "r"), ("keys", "w"), ("dirty", "w")]      # the raster kernel sorts the lists it stages, in place, and resets the key buffer it merges
STRUCTS = {
    "b2r::RasterArgs": RASTER_ARGS,
    "b2r::ShardTable": [(f"tileCount[{i}]", "r") for i in range(K_MAX_RANKS)] + [(f"tileOffset[{i}]", "r") for i in range(K_MAX_RANKS)]
                       + [(f"binIds[{i}]", "r") for i in range(K_MAX_RANKS)] + [(f"keyInbox[{i}]", "r") for i in range(K_MAX_RANKS)]
                       + [(f"dirtyInbox[{i}]", "r") for i in range(K_MAX_RANKS)],
    "b2r::ShardKeyPush": [(f"keyInbox[{i}]", "w") for i in range(K_MAX_RANKS)] + [(f"dirtyInbox[{i}]", "w") for i in range(K_MAX_RANKS)],
    "b2r::PeerFlags": [],                  # the flag words: concurrent by design, see above
    "b2r::ShardInbox": [(f"tileCount[{i}]", "w") for i in range(K_MAX_RANKS)] + [(f"tileOffset[{i}]", "w") for i in range(K_MAX_RANKS)],
    "b2r::ShardRegions": [(f"region[{i}]", "w") for i in range(K_MAX_RANKS)],
    "b2r::FrameParams": [],
}
FLAG_POINTERS = {("publish_info_kernel", 3), ("peer_barrier_kernel", 7)}       # landedFlag, timeoutFlag: flag / status words


def demangle(names):
    names = sorted(set(names))
    if not names:
        return {}
    out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True, check=True).stdout.splitlines()
    return dict(zip(names, out))


def split_params(sig):
    end = sig.rfind(")")
    depth, i = 0, end
    while i >= 0:
        depth += sig[i] == ")"
        depth -= sig[i] == "("
        if depth == 0:
            break
        i -= 1
    inner, parts, cur, d = sig[i + 1:end], [], "", 0
    for ch in inner:
        if ch in "<(":
            d += 1
        if ch in ">)":
            d -= 1
        if ch == "," and d == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        parts.append(cur.strip())
    return sig[:i], parts


def u32(raw):
    return int.from_bytes(bytes.fromhex(raw)[:4], "little")


class Trace:
    def __init__(self, path):
        self.ops = [json.loads(l) for l in open(path) if l.strip()]
        self.sig = demangle([o["kernel"] for o in self.ops if o["op"] == "launch"])
        self.parsed = {k: split_params(v) for k, v in self.sig.items()}
        self.warnings = []
        self.allocs = []          # sorted (start, end, id, kind)
        for o in self.ops:
            if o["op"] in ("malloc", "malloc_host"):
                bisect.insort(self.allocs, (o["ptr"], o["ptr"] + max(o["size"], 1), len(self.allocs), "device" if o["op"] == "malloc" else "pinned"))
        # sub-buffer boundaries: every address a kernel receives as a pointer
        self.cuts = defaultdict(set)
        for o in self.ops:
            if o["op"] == "launch":
                for _what, ptr, _mode in self.pointer_args(o)[1]:
                    a = self.find(ptr)
                    if a is not None:
                        self.cuts[a[0]].add(ptr)
        self.cuts = {k: sorted(v) for k, v in self.cuts.items()}

    def find(self, ptr):
        i = bisect.bisect_right(self.allocs, (ptr, float("inf"), 0, "")) - 1
        if i >= 0 and self.allocs[i][0] <= ptr < self.allocs[i][1]:
            return self.allocs[i]
        return None

    def short(self, o):
        return self.parsed[o["kernel"]][0].split("(")[0].replace("b2r::", "").replace("void ", "").split("<")[0]

    def pointer_args(self, o):
        """(short kernel name, [(description, pointer, mode)]) of a launch."""
        name, params = self.parsed[o["kernel"]]
        short = self.short(o)
        args = o["args"]
        if len(args) != len(params):
            self.warnings.append(f"{short}: {len(args)} logged arguments for {len(params)} parameters")
            return short, []
        acc = []
        for k, (ptype, raw) in enumerate(zip(params, args)):
            b = bytes.fromhex(raw)
            if ptype.endswith("*"):
                ptr = int.from_bytes(b[:8], "little")
                if ptr and (short, k) not in FLAG_POINTERS:
                    acc.append((f"arg {k} ({ptype})", ptr, "r" if ptype.startswith("const ") else "w"))
            elif ptype in STRUCTS:
                for j, (field, mode) in enumerate(STRUCTS[ptype]):
                    ptr = int.from_bytes(b[8 * j:8 * j + 8], "little")
                    if ptr:
                        acc.append((f"{ptype.replace('b2r::', '')}.{field}", ptr, mode))
            elif "*" in ptype or ptype.startswith("b2r::"):
                self.warnings.append(f"{short}: parameter type {ptype} is not described")
        return short, acc

    def region(self, ptr, nbytes=None):
        a = self.find(ptr)
        if a is None:
            return None
        if nbytes is None:
            cuts = self.cuts.get(a[0], [])
            i = bisect.bisect_right(cuts, ptr)
            hi = cuts[i] if i < len(cuts) else a[1]
        else:
            hi = min(ptr + nbytes, a[1])
        return (a[0], ptr, hi)

    def launch_accesses(self, o):
        short, acc = self.pointer_args(o)
        out = []
        for what, ptr, mode in acc:
            r = self.region(ptr)
            if r is None:
                self.warnings.append(f"{short}: {what} = {ptr:#x} is not inside a known allocation")
            else:
                out.append((r, mode, what))
        return short, out

    # ---- wiring: does every kernel of a frame work on that frame's buffers?
    ROLES = {      # logical buffer -> argument index, per kernel (csrc/b2r.cu: enqueueFrame / enqueueFrameRing / enqueueListSort)
        "frame_clear_kernel": {"tileCount": 0, "counters": 2, "frameInfo": 3, "hViewProj": 6, "meshStatic": 7, "meshConst": 8},
        "build_constants_kernel": {"hViewProj": 0, "meshStatic": 1, "meshConst": 2},
        "vertex_transform_kernel": {"meshConst": 2, "posIn": 3, "pos4": 4},
        "triangle_setup_kernel": {"triRefs": 0, "pos4": 1, "rects": 2, "tileCount": 3, "bigList": 4, "counters": 5, "binSlots": 6, "keys": 7, "dirty": 8, "frameInfo": 9},
        "bin_big_count_kernel": {"bigList": 0, "counters": 1, "rects": 2, "tileCount": 3},
        "tile_scan_kernel": {"tileCount": 0, "tileOffset": 1, "tileCursor": 2, "longTiles": 6, "tileOrder": 7, "dirty": 8, "counters": 11, "frameInfo": 14},
        "bin_scatter_kernel": {"rects": 0, "binSlots": 1, "tileOffset": 4, "tileCursor": 5, "binIds": 6, "bigList": 7, "counters": 8, "frameInfo": 10},      # args[12] = ShardRegions: the mirror of the lists in the owners' arenas
        "bin_sort_tiles_kernel": {"tileCount": 0, "tileOffset": 1, "binIds": 2, "frameInfo": 4},
        "bin_sort_long_kernel": {"longTiles": 0, "tileCount": 1, "tileOffset": 2, "binIds": 3, "frameInfo": 4},
        "shard_merge_heads_kernel": {"tileOrder": 4, "frameInfo": 5, "dirty": 6},
        "shard_push_keys_kernel": {"keys": 0, "dirty": 1},
        "publish_info_kernel": {"frameInfo": 0},
    }
    RASTER_ROLES = {"tileCount": 0, "tileOffset": 1, "binIds": 2, "triRefs": 3, "pos4": 4, "meshConst": 8, "frameInfo": 10, "tileOrder": 15 + 5 * K_MAX_RANKS,
                    "keys": 16 + 5 * K_MAX_RANKS, "dirty": 17 + 5 * K_MAX_RANKS}

    def frames(self):
        """Per device: the launches of every frame, in the order the host issued them (b2r_render issues a whole frame; a
        frame starts with its constants / clear kernel and ends with publish_info_kernel)."""
        out, cur = defaultdict(list), {}
        for o in self.ops:
            if o["op"] != "launch":
                continue
            dev, k = o.get("dev", 0), self.short(o)
            if k in ("build_constants_kernel", "frame_clear_kernel") and dev not in cur:
                cur[dev] = []
            if dev in cur:
                cur[dev].append(o)
                if k == "publish_info_kernel":
                    out[dev].append(cur.pop(dev))
        return out

    @staticmethod
    def _ptr(raw, index=0):
        return int.from_bytes(bytes.fromhex(raw)[8 * index:8 * index + 8], "little")

    def frame_buffers(self, frame):
        """logical buffer -> {pointer: [kernels]} over the launches of one frame."""
        seen = defaultdict(lambda: defaultdict(list))
        for o in frame:
            k = self.short(o)
            if k == "raster_shade_kernel":
                for name, j in self.RASTER_ROLES.items():
                    seen[name][self._ptr(o["args"][0], j)].append(k)
            for name, j in self.ROLES.get(k, {}).items():
                if self._ptr(o["args"][j]):                   # a null pointer = the kernel does not take part in that role
                    seen[name][self._ptr(o["args"][j])].append(k)
        return seen

    def wiring_errors(self, ranks=1):
        """Pointer-level dataflow of every frame: one buffer per logical role inside a frame (the raster kernel consumes the
        lists, positions, constants and counters its own frame produced); sharded: the peers' heads and lists a tile owner
        reads are the ones the peers pushed for that frame, and all ranks of a frame store into the same colour surface."""
        errs = []
        frames = self.frames()
        for dev, fl in frames.items():
            for i, frame in enumerate(fl):
                for name, ptrs in self.frame_buffers(frame).items():
                    if len(ptrs) > 1:
                        errs.append(f"rank {dev} frame {i}: {name} is " + " but ".join(f"{p:#x} in {sorted(set(ks))}" for p, ks in ptrs.items()))
        if ranks > 1:
            n_frames = min(len(frames[d]) for d in range(ranks))
            for i in range(n_frames):
                get = lambda d, kernel: next((o for o in frames[d][i] if self.short(o) == kernel), None)      # noqa: E731
                rasters = [get(d, "raster_shade_kernel") for d in range(ranks)]
                heads = [get(d, "tile_scan_kernel") for d in range(ranks)]          # the scan kernel pushes the heads (its ShardInbox argument)
                lists = [get(d, "bin_scatter_kernel") for d in range(ranks)]          # the scatter kernel mirrors the lists (its ShardRegions argument)
                if any(x is None for x in rasters + heads):
                    errs.append(f"frame {i}: a rank has no raster / push-heads launch")
                    continue
                # the ranks' primitive slices: contiguous, in rank order, no gap and no overlap (a rank with an empty slice
                # launches no setup kernel); the scatter kernel works on the slice the setup kernel binned
                slices = []
                for d in range(ranks):
                    setup, scatter = get(d, "triangle_setup_kernel"), get(d, "bin_scatter_kernel")
                    if setup is not None:
                        lo, hi = u32(setup["args"][10]), u32(setup["args"][11])
                        slices.append((lo, hi, d))
                        if scatter is not None and (u32(scatter["args"][2]), u32(scatter["args"][3])) != (lo, hi):
                            errs.append(f"frame {i}: rank {d} bins primitives [{lo}, {hi}) but scatters [{u32(scatter['args'][2])}, {u32(scatter['args'][3])})")
                if slices:
                    if slices[0][0] != 0 or any(a[1] != b[0] for a, b in zip(slices, slices[1:])) or any(lo >= hi for lo, hi, _ in slices):
                        errs.append(f"frame {i}: the ranks' primitive slices {[(lo, hi) for lo, hi, _ in slices]} do not tile the primitives in rank order")
                # the flag epochs of a frame agree on all ranks
                epochs = {tuple((u32(o["args"][3]), u32(o["args"][4])) for o in frames[d][i] if self.short(o) == "peer_barrier_kernel" and u32(o["args"][5]))
                          for d in range(ranks)}
                if len(epochs) != 1:
                    errs.append(f"frame {i}: the ranks' barrier epochs differ: {sorted(epochs)}")
                surfaces = {self._ptr(r["args"][0], 11) for r in rasters}
                if len(surfaces) != 1:
                    errs.append(f"frame {i}: the ranks store their colours into different surfaces {sorted(hex(x) for x in surfaces)}")
                for q in range(ranks):                     # tile owner q reads ...
                    for p in range(ranks):                 # ... what rank p pushed
                        base = 15                          # RasterArgs.shard: tileCount[8], tileOffset[8], binIds[8]
                        want_count, want_off = self._ptr(heads[p]["args"][12], q), self._ptr(heads[p]["args"][12], K_MAX_RANKS + q)
                        got_count, got_off = self._ptr(rasters[q]["args"][0], base + p), self._ptr(rasters[q]["args"][0], base + K_MAX_RANKS + p)
                        if (want_count, want_off) != (got_count, got_off):
                            errs.append(f"frame {i}: rank {q} reads the heads of rank {p} at {got_count:#x}/{got_off:#x}, rank {p} pushed them to {want_count:#x}/{want_off:#x}")
                        got_ids = self._ptr(rasters[q]["args"][0], base + 2 * K_MAX_RANKS + p)
                        if p == q:
                            want_ids = self._ptr(rasters[q]["args"][0], 2)
                        else:
                            want_ids = self._ptr(lists[p]["args"][12], q) if lists[p] is not None else got_ids
                        if want_ids != got_ids:
                            errs.append(f"frame {i}: rank {q} reads the lists of rank {p} at {got_ids:#x}, they are at {want_ids:#x}")
        return errs

    def buffer_sets(self, dev=0, name="tileCount"):
        """The sequence of pointers a per-frame buffer takes over the frames of one device (ring slots / frame parity)."""
        return [next(iter(self.frame_buffers(f)[name])) for f in self.frames()[dev]]

    def api_errors(self):
        """Calls the recording runtime refused because a real device would (e.g. elapsed time between events without timing)."""
        return [o["what"] for o in self.ops if o["op"] == "api_error"]

    def leaks(self):
        """Device and pinned allocations that are never released (the scenario destroys its contexts before it ends)."""
        live = {}
        for o in self.ops:
            if o["op"] in ("malloc", "malloc_host"):
                live[o["ptr"]] = (o["op"], o["size"])
            elif o["op"] in ("free", "free_host"):
                live.pop(o["ptr"], None)
        return live

    def bad_launches(self):
        """Launch configurations a real device rejects (cudaErrorInvalidConfiguration) but the recording runtime accepts."""
        bad = []
        for o in self.ops:
            if o["op"] == "launch":
                g, b = o["grid"], o["block"]
                if min(g) < 1 or min(b) < 1 or g[0] > 2**31 - 1 or g[1] > 65535 or g[2] > 65535 or b[0] * b[1] * b[2] > 1024:
                    bad.append(f"{self.short(o)} grid {g} block {b}")
        return bad

    # ---- the graph
    def analyse(self):
        nodes = []                                   # {label, idx, dev, acc, preds}
        last_on_stream, last_record, host = {}, {}, {}
        arrive = defaultdict(list)                   # (slot base, epoch) -> arrive nodes of the barrier kernels
        complete = defaultdict(list)
        signals = defaultdict(list)                  # (slot base, epoch) -> signalling nodes (publish_info_kernel)
        waiters = defaultdict(list)

        def node(label, preds, dev, acc=(), idx=None):
            nodes.append({"label": label, "idx": idx, "dev": dev, "acc": list(acc), "preds": [p for p in preds if p is not None]})
            return len(nodes) - 1

        def host_node(label, preds, dev):
            host[dev] = node("host: " + label, [host.get(dev)] + list(preds), dev)

        for idx, o in enumerate(self.ops):
            op, dev = o["op"], o.get("dev", 0)
            skey = (dev, 0) if o.get("stream", 1) == 0 else o.get("stream")
            if op == "host_barrier":
                # the ranks' host threads meet (torch.distributed.barrier in a real group): what any of them has waited for
                # precedes everything any of them issues afterwards
                n = node("host barrier between the ranks", list(host.values()), dev)
                for d in {nd["dev"] for nd in nodes} | set(host):
                    host[d] = n
            elif op == "free":
                host_node("cudaFree", [n for k, n in last_on_stream.items() if nodes[n]["dev"] == dev], dev)     # synchronises the device
            elif op == "stream_sync":
                host_node(f"sync stream {o['stream']}", [last_on_stream.get(skey)], dev)
            elif op == "event_sync":
                host_node(f"sync event {o['event']}", [last_record.get(o["event"])], dev)
            elif op == "event_record":
                n = node(f"record e{o['event']} on s{o['stream']}", [last_on_stream.get(skey), host.get(dev)], dev, idx=idx)
                last_on_stream[skey] = n
                last_record[o["event"]] = n
            elif op == "stream_wait":
                n = node(f"s{o['stream']} waits e{o['event']}", [last_on_stream.get(skey), last_record.get(o["event"]), host.get(dev)], dev, idx=idx)
                last_on_stream[skey] = n
            elif op in ("launch", "memcpy", "memset"):
                if op == "launch":
                    label, acc = self.launch_accesses(o)
                elif op == "memcpy":
                    label, acc = f"memcpy {o['bytes']} B", []
                    for ptr, mode in ((o["dst"], "w"), (o["src"], "r")):
                        r = self.region(ptr, o["bytes"])
                        if r is not None and o["bytes"]:
                            acc.append((r, mode, "dst" if mode == "w" else "src"))
                else:
                    label, acc = f"memset {o['bytes']} B", []
                    r = self.region(o["dst"], o["bytes"])
                    if r is not None:
                        acc.append((r, "w", "dst"))
                tag = f"{label} on s{o['stream']}" + (f" (rank {dev})" if dev else "")
                if op == "launch" and label == "peer_barrier_kernel":
                    a = o["args"]
                    base, epoch, signal, wait = u32(a[3]), u32(a[4]), u32(a[5]), u32(a[6])
                    n_arrive = node(tag + f" [arrive {base}/{epoch}]", [last_on_stream.get(skey), host.get(dev)], dev, idx=idx)
                    n = node(tag + f" [pass {base}/{epoch}]", [n_arrive], dev, idx=idx)
                    if signal:
                        arrive[(base, epoch)].append(n_arrive)
                        complete[(base, epoch)].append(n)
                    elif wait:
                        waiters[(base, epoch)].append(n)
                else:
                    n = node(tag, [last_on_stream.get(skey), host.get(dev)], dev, acc, idx)
                    if op == "launch" and label == "publish_info_kernel":
                        a = o["args"]
                        landed, epoch, done_ranks = int.from_bytes(bytes.fromhex(a[3])[:8], "little"), u32(a[4]), u32(a[6])
                        if landed:
                            signals[(8, epoch)].append(n)
                        if done_ranks:
                            signals[(16, epoch)].append(n)
                    if op == "launch" and label == "frame_clear_kernel" and len(o["args"]) > 13 and u32(o["args"][12]):
                        # frame ring + sharded front end: block 0 of the clear kernel ends with the one-way wait for the
                        # peers' "frame done" of the frame that used this ring slot before (wait ranks, wait epoch)
                        waiters[(16, u32(o["args"][13]))].append(n)
                last_on_stream[skey] = n
                if op == "memcpy" and not o["async"]:
                    host_node("blocking copy", [n], dev)
        for key, cs in complete.items():             # barrier: every arrival precedes every completion
            for c in cs:
                nodes[c]["preds"] += arrive[key]
        for key, ws in waiters.items():              # one-way waits: every rank's signal of that epoch precedes the waiter
            for w in ws:
                nodes[w]["preds"] += signals.get(key, [])
        self.nodes = nodes
        self.reach()
        return self.races()

    def reach(self):
        n = len(self.nodes)
        succ = [[] for _ in range(n)]
        indeg = [0] * n
        for j, nd in enumerate(self.nodes):
            for p in set(nd["preds"]):
                succ[p].append(j)
                indeg[j] += 1
        order, stack = [], [i for i in range(n) if indeg[i] == 0]
        while stack:
            i = stack.pop()
            order.append(i)
            for j in succ[i]:
                indeg[j] -= 1
                if indeg[j] == 0:
                    stack.append(j)
        if len(order) != n:
            raise RuntimeError("the ordering constraints contain a cycle: the schedule would deadlock")
        anc = [0] * n
        for j in order:
            bits = 0
            for p in self.nodes[j]["preds"]:
                bits |= anc[p] | (1 << p)
            anc[j] = bits
        self.anc = anc

    def races(self):
        by_alloc = defaultdict(list)
        for n, nd in enumerate(self.nodes):
            for (alloc, lo, hi), mode, what in nd["acc"]:
                by_alloc[alloc].append((n, lo, hi, mode, what))
        found = []
        for alloc, accs in by_alloc.items():
            for j in range(len(accs)):
                nj, loj, hij, mj, wj = accs[j]
                for i in range(j):
                    ni, loi, hii, mi, wi = accs[i]
                    if ni == nj or (mi == "r" and mj == "r") or hii <= loj or hij <= loi:
                        continue
                    # sharded front end: all ranks store their tiles' colours into rank 0's surface; tile ownership is
                    # static (tile index mod ranks), so two ranks never write the same pixel, in whatever frames
                    if wi == wj == "RasterArgs.framebuffer" and self.nodes[ni]["dev"] != self.nodes[nj]["dev"]:
                        continue
                    if not ((self.anc[nj] >> ni) & 1 or (self.anc[ni] >> nj) & 1):
                        found.append((ni, nj, alloc, wi, wj, mi, mj))
        return found

    def describe(self, race):
        ni, nj, alloc, wi, wj, mi, mj = race
        a = self.find(alloc)
        return (f"{self.nodes[ni]['label']} [log line {self.nodes[ni]['idx'] + 1}] {'writes' if mi == 'w' else 'reads'} {wi}  ||  "
                f"{self.nodes[nj]['label']} [log line {self.nodes[nj]['idx'] + 1}] {'writes' if mj == 'w' else 'reads'} {wj}   "
                f"(allocation #{a[2]}, {a[1] - a[0]} bytes, {a[3]})")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("log")
    ap.add_argument("-v", action="store_true")
    a = ap.parse_args()
    t = Trace(a.log)
    races = t.analyse()
    launches = sum(1 for o in t.ops if o["op"] == "launch")
    print(f"{len(t.ops)} logged operations, {launches} kernel launches, {len(t.nodes)} graph nodes, {len(t.allocs)} allocations, {len(races)} unordered conflicting pairs")
    for w in sorted(set(t.warnings)):
        print("warning:", w)
    for b in t.bad_launches():
        print("INVALID LAUNCH:", b)
    ranks = len({o.get("dev", 0) for o in t.ops if o["op"] == "launch"})
    wiring = t.wiring_errors(ranks)
    for w in wiring[:20]:
        print("WIRING:", w)
    seen = set()
    for r in races:
        key = (t.nodes[r[0]]["label"].split(" on ")[0], t.nodes[r[1]]["label"].split(" on ")[0], r[3], r[4])
        if key in seen and not a.v:
            continue
        seen.add(key)
        print("RACE:", t.describe(r))
    return 1 if races or t.bad_launches() or wiring else 0


if __name__ == "__main__":
    sys.exit(main())
