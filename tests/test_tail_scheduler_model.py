"""Executable model of the job protocol of the persistence launch (csrc/ph_small.cuh: the job loop, `map_matched`, and the
tail's scheduler), run under RANDOM interleavings of the SMs at the granularity of their atomic operations.

What the kernel relies on, and what this checks for every schedule:
  * termination: no SM waits for ever (a held gradient job waits for matching jobs, which wait for persistence jobs, which
    wait for nothing; an SM that holds a gradient job keeps claiming and running matching jobs);
  * every map is matched exactly once -- where its second diagram completes when one of the two diagrams is empty
    (ready = 4), otherwise by exactly one matching job of the tail (ready = 3 -> claimed by fetch-add);
  * every image is published exactly once, by whoever matches its last map, and never before all its maps have a cost;
  * every (image, channel) of a published image gets exactly one gradient job, run only after the publication; images
    with a map on the heavy list get none (they are left to grad_kernel: gfused stays 0).

The CUDA kernels themselves are checked against the oracle in test_gpu_parity.py (incl. the fused gradient against the
separate launch); compute-sanitizer's racecheck is not available on the GPU pool, so this model is the protocol's
CPU-side evidence next to the repeatability tests.
"""
import random

import pytest

VALID, SKIP = 1 << 31, 1 << 30


class World:
    def __init__(self, B, C, rng, p_empty, p_heavy, fuse_grad=True):
        self.B, self.C, self.M = B, C, B * C
        self.n_jobs = 2 * self.M                      # all prediction maps first, then the ground-truth maps
        self.job_counter = 0
        self.ready = [0] * self.M
        # diagram sizes as the persistence jobs will write them: an empty truth diagram is the common case
        self.n_pred = [rng.choice([0, 3, 50]) if rng.random() < 0.2 else 40 for _ in range(self.M)]
        self.n_true = [0 if rng.random() < p_empty else rng.choice([1, 2, 5]) for _ in range(self.M)]
        self.heavy = [rng.random() < p_heavy for _ in range(self.M)]   # matching job hands the map to the general kernel
        self.diagram_done = [[False, False] for _ in range(self.M)]
        self.m_counter = 0
        self.img_cnt = [0] * B
        self.gq = [0] * B
        self.gq_tail = 0
        self.gq_head = 0
        self.gfused = [0] * B
        self.fuse_grad = fuse_grad
        # bookkeeping of the checks
        self.cost_written = [0] * self.M
        self.matching_jobs = [0] * self.M
        self.grad_jobs = {}
        self.published = [0] * B


def map_matched(w, k, heavy):
    """ph_small.cuh `map_matched` (thread 0): count the map for its image; the last one publishes the image."""
    b = k // w.C
    old = w.img_cnt[b]
    w.img_cnt[b] += 0x10001 if heavy else 1      # one atomicAdd
    yield
    if (old & 0xFFFF) + 1 == w.C:
        skip = (old >> 16) != 0 or heavy
        if not skip:
            w.gfused[b] = 1
        slot = w.gq_tail
        w.gq_tail += 1                           # atomicAdd
        yield
        assert all(w.cost_written[b * w.C + c] == 1 for c in range(w.C)), "image published before all its maps have a cost"
        w.published[b] += 1
        w.gq[slot] = VALID | (SKIP if skip else 0) | b   # st.release


def sm(w, rng):
    """One CTA: the persistence job loop, then the tail's scheduler."""
    # ---- persistence jobs (claimed one ahead in the kernel; the claim order is what matters here)
    while True:
        job = w.job_counter
        w.job_counter += 1                       # atomicAdd
        yield
        if job >= w.n_jobs:
            break
        s, k = divmod(job, w.M)
        for _ in range(rng.randint(1, 4 if s == 0 else 2)):   # the map's phases
            yield
        w.diagram_done[k][s] = True
        old = w.ready[k]
        w.ready[k] += 1                          # atomicAdd(ready + k, 1)
        yield
        if old == 1:                             # this SM finished the map's second diagram
            assert all(w.diagram_done[k])
            st = 3
            if w.n_pred[k] == 0 or w.n_true[k] == 0:   # cost = the two diagonal sums: written here
                w.cost_written[k] += 1
                st = 4
                if w.fuse_grad:
                    yield from map_matched(w, k, False)
            yield
            w.ready[k] = st                      # st.release
    # ---- tail
    n_gjobs = w.M if w.fuse_grad else 0
    held_m = held_g = None
    m_done, g_done = False, n_gjobs == 0
    held_m = w.m_counter
    w.m_counter += 1
    yield
    if not g_done:
        held_g = w.gq_head
        w.gq_head += 1
        yield
    while True:
        kind = arg = 0
        while True:
            if not m_done and held_m is None:
                held_m = w.m_counter
                w.m_counter += 1
                yield
            if not g_done and held_g is None:
                held_g = w.gq_head
                w.gq_head += 1
                yield
            if not m_done and held_m >= w.M:
                m_done, held_m = True, None
            if not g_done and held_g >= n_gjobs:
                g_done, held_g = True, None
            if held_m is not None:
                v = w.ready[held_m]              # ld.acquire
                yield
                if v >= 4:
                    held_m = None
                    continue
                if v == 3:
                    kind, arg, held_m = 1, held_m, None
                    break
            if held_g is not None:
                e = w.gq[held_g // w.C]          # ld.acquire
                yield
                if e & VALID:
                    ch = held_g % w.C
                    held_g = None
                    if e & SKIP:
                        continue
                    kind, arg = 2, (e & 0xFFFFFF) * w.C + ch
                    break
            if m_done and g_done:
                break
            yield                                # __nanosleep
        if kind == 1 and not m_done:             # claim ahead
            held_m = w.m_counter
            w.m_counter += 1
            yield
        if kind == 2 and not g_done:
            held_g = w.gq_head
            w.gq_head += 1
            yield
        if kind == 0:
            return
        if kind == 1:
            w.matching_jobs[arg] += 1
            for _ in range(rng.randint(1, 3)):
                yield
            w.cost_written[arg] += 1             # match_one_map wrote the cost (or put the map on the heavy list)
            if w.fuse_grad:
                yield from map_matched(w, arg, w.heavy[arg])
        else:
            b = arg // w.C
            assert w.published[b] == 1 and all(w.cost_written[b * w.C + c] == 1 for c in range(w.C))
            w.grad_jobs[arg] = w.grad_jobs.get(arg, 0) + 1
            for _ in range(rng.randint(1, 3)):
                yield


def run(B, C, n_sm, seed, p_empty, p_heavy, fuse_grad=True):
    rng = random.Random(seed)
    w = World(B, C, rng, p_empty, p_heavy, fuse_grad)
    agents = [sm(w, rng) for _ in range(n_sm)]
    live = list(range(n_sm))
    steps, limit = 0, 200 * (w.n_jobs + n_sm) * 20
    # a schedule may starve SMs for a while: pick with a random bias that changes over time
    while live:
        steps += 1
        assert steps < limit, "the protocol did not terminate under this schedule"
        if rng.random() < 0.02:
            rng.shuffle(live)
        i = live[min(int(rng.expovariate(0.6)), len(live) - 1)]
        try:
            next(agents[i])
        except StopIteration:
            live.remove(i)
    return w


@pytest.mark.parametrize("seed", range(40))
def test_tail_protocol_under_random_schedules(seed):
    rng = random.Random(1000 + seed)
    B, C = rng.randint(1, 6), rng.randint(1, 5)
    n_sm = rng.choice([1, 2, 3, 7, 16])
    p_empty, p_heavy = rng.choice([0.0, 0.5, 0.9, 1.0]), rng.choice([0.0, 0.1, 0.5])
    w = run(B, C, n_sm, seed, p_empty, p_heavy)
    for k in range(w.M):
        trivial = w.n_pred[k] == 0 or w.n_true[k] == 0
        assert w.cost_written[k] == 1, (k, "matched exactly once")
        assert w.matching_jobs[k] == (0 if trivial else 1)
        assert w.ready[k] == (4 if trivial else 3)
    assert w.gq_tail == B and w.published == [1] * B
    seen = sorted(e & 0xFFFFFF for e in w.gq)
    assert seen == list(range(B)) and all(e & VALID for e in w.gq)
    for b in range(B):
        has_heavy = any(w.heavy[b * C + c] and not (w.n_pred[b * C + c] == 0 or w.n_true[b * C + c] == 0) for c in range(C))
        assert w.gfused[b] == (0 if has_heavy else 1)
        for c in range(C):
            assert w.grad_jobs.get(b * C + c, 0) == (0 if has_heavy else 1), (b, c)


@pytest.mark.parametrize("seed", range(8))
def test_tail_protocol_without_the_fused_gradient(seed):
    """tl_forward: matching jobs only (no publication queue)."""
    w = run(3, 4, 5, seed, 0.5, 0.2, fuse_grad=False)
    assert all(c == 1 for c in w.cost_written) and not w.grad_jobs and w.gq_tail == 0
