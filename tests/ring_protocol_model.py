"""Timing model of the frame ring of the sharded front end (csrc/b2r.cu:enqueueFrameRing), used by the CPU tests.

It replays the ORDERING RULES the host code enqueues - stream order inside a ring stream, the event that chains the
raster kernels, the three kinds of flag words with the epochs the kernels store and wait for - under random kernel
durations and rank skew, and reports every access to a list set that the rules fail to order:

  S1  rank r's stores of frame f into rank p's set f % D (list heads after the scan, lists after the sort) must not
      begin before p's raster kernel of frame f - D (the last reader of that set) has finished;
  S2  p's raster kernel of frame f must not begin before every rank's lists of frame f have landed;
  S3  the epochs a rank publishes ("frame done" / "colours landed") must rise monotonically.

Per frame f and rank r, on ring stream f % D (a later frame of the same stream starts after this one has ended):

  front_a    clear, vertex, setup, scan                     (local set only)
  done_wait  f >= D: until done[r][p] >= f - D + 1 for every p   (flag words 16..23, epoch = frames finished)
  push       list heads -> peers, scatter, sort, lists -> peers  (stores into set f % D of every peer)
  barrier    lists[p][r] = f + 1 for every p; until lists[r][p] >= f + 1 for every p   (flag words 0..7)
  raster     after the event of this rank's frame f - 1 (raster + publish); reads set f % D of its own arena
  publish    done[p][r] = f + 1 for every p
"""
from __future__ import annotations

import numpy as np


def simulate(n_ranks: int, depth: int, n_frames: int, rng: np.random.Generator, *, done_wait: bool = True, done_lag: int = 0,
             chain_rasters: bool = True, skew: float = 3.0):
    """Returns (violations, frame_times).  `done_wait=False` drops the wait for the peers' "frame done" flags,
    `done_lag=1` waits for the frame before the right one (an off-by-one in the epoch), `chain_rasters=False` drops
    the event between consecutive raster kernels - the negative controls of the test."""
    R, D, F = n_ranks, depth, n_frames
    speed = 1.0 + skew * rng.random(R)                       # some ranks are persistently slower
    dur = lambda r, lo, hi: float(speed[r] * rng.uniform(lo, hi))
    stream_free = np.zeros((R, D))                           # when ring stream k of rank r has finished its newest frame
    push_begin, push_end = np.zeros((F, R)), np.zeros((F, R))
    raster_begin, raster_end, publish_t = np.zeros((F, R)), np.zeros((F, R)), np.zeros((F, R))
    host_t = np.zeros(R)                                     # the host enqueues frames in order, a little apart
    violations = []
    for f in range(F):
        k = f % D
        signal_t = np.zeros(R)
        for r in range(R):
            host_t[r] += dur(r, 0.0, 2.0)
            t = max(stream_free[r, k], host_t[r])
            t += dur(r, 5.0, 40.0)                           # front_a
            if done_wait and f >= D:
                need = f - D + 1 - done_lag                  # epoch the kernel waits for
                if need >= 1:
                    # done[r][p] reaches `need` when rank p publishes frame need - 1
                    t = max(t, max(publish_t[need - 1, p] for p in range(R)))
            push_begin[f, r] = t
            t += dur(r, 5.0, 60.0)                           # heads -> scatter -> sort -> lists
            push_end[f, r] = t
            signal_t[r] = t
        ready = float(signal_t.max())                        # every rank's lists[.][p] >= f + 1
        for r in range(R):
            t = max(signal_t[r], ready) + dur(r, 0.1, 1.0)   # barrier kernel
            if chain_rasters and f > 0:
                t = max(t, publish_t[f - 1, r])
            raster_begin[f, r] = t
            t += dur(r, 5.0, 80.0)
            raster_end[f, r] = t
            t += dur(r, 0.1, 1.0)
            publish_t[f, r] = t
            stream_free[r, k] = t
        # ---- checks
        for r in range(R):
            for p in range(R):
                if f >= D and push_begin[f, r] < raster_end[f - D, p]:
                    violations.append(("S1", f, r, p))
                if raster_begin[f, p] < push_end[f, r]:
                    violations.append(("S2", f, r, p))
            if f > 0 and publish_t[f, r] < publish_t[f - 1, r]:
                violations.append(("S3", f, r, r))
    return violations, publish_t.max(axis=1)
