#!/usr/bin/env python
"""Rewrites the CUDA-only SYNTAX of the product's sources into calls of tools/emu/emu_cuda.h, into build/emu/src (scratch):
  kernel<<<grid, block[, smem[, stream]]>>>(args)  ->  emu::launch("kernel", kernel, dim3(grid), dim3(block), smem, stream, args)
  cudaLaunchKernelEx(                               ->  emu::launchEx(
  asm volatile(                                     ->  EMU_ASM(
  cudaFuncSetAttribute(                             ->  emu::funcSetAttribute(
  asm("rcp.approx.ftz.f32 ...")                     ->  r = 1.0f / b;
  asm volatile("cp.async.ca.shared.global ...")     ->  memcpy(smemDst, gsrc, 16);
  extern __shared__ T name[];                       ->  T* name = (T*)emu::dynamicSmemPtr();
Nothing else is touched: the kernels that run in the emulator are the product's kernels."""
import os
import re
import sys


def split_top(s):
    parts, cur, d = [], "", 0
    for ch in s:
        if ch in "([{":
            d += 1
        if ch in ")]}":
            d -= 1
        if ch == "," and d == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
    parts.append(cur.strip())
    return parts


def rewrite_launches(src):
    out, i = "", 0
    while True:
        j = src.find("<<<", i)
        if j < 0:
            return out + src[i:]
        m = re.search(r"([A-Za-z_][A-Za-z0-9_:]*)\s*$", src[i:j])
        name = m.group(1)
        out += src[i:i + m.start(1)]
        k = src.find(">>>", j)
        cfg = split_top(src[j + 3:k])
        a = src.index("(", k)
        d, b = 0, a
        while True:
            d += src[b] == "("
            d -= src[b] == ")"
            if d == 0:
                break
            b += 1
        args = src[a + 1:b].strip()
        smem = cfg[2] if len(cfg) > 2 else "0"
        stream = cfg[3] if len(cfg) > 3 else "nullptr"
        out += f'emu::launch("{name}", {name}, dim3({cfg[0]}), dim3({cfg[1]}), {smem}, (cudaStream_t)({stream})' + (", " + args if args else "") + ")"
        i = b + 1


def main(src_dir, dst_dir):
    os.makedirs(os.path.join(dst_dir, "host"), exist_ok=True)
    for rel in ("b2r.cu", "device_types.cuh", "geometry_kernels.cuh", "raster_kernel.cuh"):
        s = open(os.path.join(src_dir, rel)).read()
        s = rewrite_launches(s)
        s = s.replace("cudaLaunchKernelEx(", "emu::launchEx(")
        s = s.replace('asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));', "t = emuGlobalTimerNs();")
        # cp.async: a plain copy at issue time (the emulator has no asynchronous copies; commit / wait become no-ops below)
        s = s.replace('asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"((uint32_t)__cvta_generic_to_shared(smemDst)), "l"(gsrc) : "memory");', "memcpy(smemDst, gsrc, 16);")
        # ... and its form with a converted shared-memory address: the "address" is the host pointer
        s = s.replace('asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(smemDst), "l"(gsrc) : "memory");', "memcpy((void*)smemDst, gsrc, 16);")
        s = s.replace("typedef uint32_t SmemAddr;", "typedef uintptr_t SmemAddr;")
        s = s.replace("return (SmemAddr)__cvta_generic_to_shared(p);", "return (SmemAddr)p;")
        s = s.replace("asm volatile(", "EMU_ASM(")
        s = s.replace("cudaFuncSetAttribute(", "emu::funcSetAttribute(")
        # MUFU.RCP: the correctly rounded reciprocal stands in for the hardware approximation; the refinement steps that follow
        # it in the product (device_types.cuh: refinedReciprocal / divByArea) converge to the same correctly rounded quotient
        s = s.replace('asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));', "r = 1.0f / b;")
        s = re.sub(r"extern\s+__shared__\s+(\w+)\s+(\w+)\[\];", r"\1* \2 = (\1*)emu::dynamicSmemPtr();", s)
        open(os.path.join(dst_dir, rel.replace(".cu", ".cpp") if rel.endswith(".cu") else rel), "w").write(s)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
