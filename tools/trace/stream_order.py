#!/usr/bin/env python
"""Stream-order race detector for the frame scheduler - needs no GPU.

Input: the JSON-lines log the recording CUDA runtime (tools/trace/fake_cudart.cpp) writes while the product's own host
code (build/trace/libb2r_trace.so) runs.  Every device operation (kernel launch, copy, memset) becomes a node with the
memory ranges it reads / writes; edges are what CUDA guarantees: order within a stream, event record -> stream wait,
and host synchronisation (cudaStreamSynchronize / cudaEventSynchronize / cudaFree / blocking copies: everything the host
has waited for precedes everything it issues afterwards).  Two operations CONFLICT when they touch overlapping bytes of
an allocation, at least one of them writes, and neither is ordered before the other: that is a data race between
frames in flight, whatever the timing on a real GPU happens to be.

What a kernel touches is derived from its launch arguments: pointer parameters from the demangled signature
(`const T*` = read, `T*` = write), the by-value structs RasterArgs / PeerFlags / ShardInbox / ShardRegions from their
layout (csrc/raster_kernel.cuh, csrc/device_types.cuh).  A pointer argument stands for the range from the pointer to the
end of its allocation, except the colour surface of the raster kernel (two surfaces share one allocation:
width * height * 4 bytes).  Programmatic dependent launch does not relax stream order here: every kernel of the product
begins with griddepcontrol.wait (pdlPrologue), i.e. waits for the completion and memory flush of its predecessor.

    python tools/trace/stream_order.py log.jsonl --width W --height H [-v]
exit status 1 when a race is found.  tests/test_cpu_trace.py drives it."""
from __future__ import annotations

import argparse
import bisect
import json
import re
import subprocess
import sys
from collections import defaultdict

K_MAX_RANKS = 8
RASTER_ARGS = [("tileCount", "r"), ("tileOffset", "r"), ("binIds", "r"), ("triRefs", "r"), ("pos4", "r"), ("posIn", "r"), ("ivertex", "r"),
               ("meshes", "r"), ("textures", "r"), ("frameInfo", "w"), ("framebuffer", "w"), ("depth", "w"), ("zSurface", "w"), ("winnerSurface", "w")]
RASTER_ARGS += [(f"shard.tileCount[{i}]", "r") for i in range(K_MAX_RANKS)] + [(f"shard.tileOffset[{i}]", "r") for i in range(K_MAX_RANKS)] \
    + [(f"shard.binIds[{i}]", "r") for i in range(K_MAX_RANKS)]
STRUCTS = {
    "b2r::RasterArgs": RASTER_ARGS,
    "b2r::PeerFlags": [(f"flags[{i}]", "w") for i in range(K_MAX_RANKS)],
    "b2r::ShardInbox": [(f"tileCount[{i}]", "w") for i in range(K_MAX_RANKS)] + [(f"tileOffset[{i}]", "w") for i in range(K_MAX_RANKS)],
    "b2r::ShardRegions": [(f"region[{i}]", "w") for i in range(K_MAX_RANKS)],
    "b2r::FrameParams": [],
}


def demangle(names):
    names = sorted(set(names))
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


class Trace:
    def __init__(self, path, width, height):
        self.ops = [json.loads(l) for l in open(path) if l.strip()]
        self.width, self.height = width, height
        self.allocs = []          # sorted list of (start, end, id, kind)
        self.sig = demangle([o["kernel"] for o in self.ops if o["op"] == "launch"])
        self.warnings = []

    # ---- memory map
    def add_alloc(self, ptr, size, kind):
        bisect.insort(self.allocs, (ptr, ptr + max(size, 1), len(self.allocs), kind))

    def find(self, ptr):
        i = bisect.bisect_right(self.allocs, (ptr, float("inf"), 0, "")) - 1
        if i >= 0 and self.allocs[i][0] <= ptr < self.allocs[i][1]:
            return self.allocs[i]
        return None

    def region(self, ptr, nbytes=None):
        a = self.find(ptr)
        if a is None:
            return None
        return (a[0], ptr, a[1] if nbytes is None else min(ptr + nbytes, a[1]))      # (allocation, lo, hi)

    # ---- accesses of one operation
    def launch_accesses(self, o):
        name, params = split_params(self.sig[o["kernel"]])
        short = name.split("(")[0].replace("b2r::", "").replace("void ", "")
        args = o["args"]
        if len(args) != len(params):
            self.warnings.append(f"{short}: {len(args)} logged arguments for {len(params)} parameters")
            return short, []
        acc = []
        for ptype, raw in zip(params, args):
            b = bytes.fromhex(raw)
            if ptype.endswith("*"):
                ptr = int.from_bytes(b[:8], "little")
                if ptr:
                    acc.append((ptype, ptr, "r" if ptype.startswith("const ") else "w", None))
            elif ptype in STRUCTS:
                for k, (field, mode) in enumerate(STRUCTS[ptype]):
                    ptr = int.from_bytes(b[8 * k:8 * k + 8], "little")
                    if ptr:
                        acc.append((f"{ptype}.{field}", ptr, mode, self.width * self.height * 4 if field == "framebuffer" else None))
            elif "*" in ptype or ptype.startswith("b2r::"):
                self.warnings.append(f"{short}: parameter type {ptype} is not described")
        out = []
        for what, ptr, mode, nbytes in acc:
            r = self.region(ptr, nbytes)
            if r is None:
                self.warnings.append(f"{short}: {what} = {ptr:#x} is not inside a known allocation")
            else:
                out.append((r, mode, what))
        return short, out

    # ---- the graph
    def analyse(self):
        nodes = []                # dicts: label, issue index, accesses [(region, mode, what)]
        anc = []                  # ancestor bitsets
        last_on_stream, last_record, host = {}, {}, None

        def node(label, preds, accesses=(), idx=None):
            bits = 0
            for p in preds:
                if p is not None:
                    bits |= anc[p] | (1 << p)
            nodes.append({"label": label, "acc": list(accesses), "idx": idx})
            anc.append(bits)
            return len(nodes) - 1

        def host_node(label, preds):
            nonlocal host
            host = node("host: " + label, [host] + list(preds))

        for idx, o in enumerate(self.ops):
            op = o["op"]
            if op == "malloc":
                self.add_alloc(o["ptr"], o["size"], "device")
            elif op == "malloc_host":
                self.add_alloc(o["ptr"], o["size"], "pinned")
            elif op == "free":
                host_node("cudaFree", list(last_on_stream.values()))            # cudaFree synchronises the device
            elif op == "stream_sync":
                host_node(f"sync stream {o['stream']}", [last_on_stream.get(o["stream"])])
            elif op == "event_sync":
                host_node(f"sync event {o['event']}", [last_record.get(o["event"])])
            elif op == "event_record":
                n = node(f"record e{o['event']} on s{o['stream']}", [last_on_stream.get(o["stream"]), host], idx=idx)
                last_on_stream[o["stream"]] = n
                last_record[o["event"]] = n
            elif op == "stream_wait":
                n = node(f"s{o['stream']} waits e{o['event']}", [last_on_stream.get(o["stream"]), last_record.get(o["event"]), host], idx=idx)
                last_on_stream[o["stream"]] = n
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
                s = o["stream"]
                blocking = op == "memcpy" and not o["async"]
                # stream 0 = the legacy default stream: ordered in itself, but it does not order anything against the
                # product's non-blocking streams
                n = node(f"{label} on s{s}", [last_on_stream.get(s), host], acc, idx)
                last_on_stream[s] = n
                if blocking:
                    host_node("blocking copy", [n])
        self.nodes, self.anc = nodes, anc
        return self.races()

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
                    if not (self.anc[nj] >> ni) & 1:
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
    ap.add_argument("--width", type=int, required=True)
    ap.add_argument("--height", type=int, required=True)
    ap.add_argument("-v", action="store_true")
    a = ap.parse_args()
    t = Trace(a.log, a.width, a.height)
    races = t.analyse()
    launches = sum(1 for o in t.ops if o["op"] == "launch")
    print(f"{len(t.ops)} logged operations, {launches} kernel launches, {len(t.nodes)} graph nodes, {len(t.allocs)} allocations, {len(races)} unordered conflicting pairs")
    for w in sorted(set(t.warnings)):
        print("warning:", w)
    seen = set()
    for r in races:
        key = (t.nodes[r[0]]["label"].split(" on ")[0], t.nodes[r[1]]["label"].split(" on ")[0], r[3], r[4])
        if key in seen and not a.v:
            continue
        seen.add(key)
        print("RACE:", t.describe(r))
    return 1 if races else 0


if __name__ == "__main__":
    sys.exit(main())
