"""Deadlock / ordering check of the mbarrier protocol of `gemm_tc_bf16_kernel` (hg_gemm_tc.cu), offline.

Each role of the kernel (TMA producer, 8 splitter warps, MMA issuer, 4 epilogue warps) is a Python generator that
mirrors the kernel's loop structure and yields ("wait", barrier, parity) / ("arrive", barrier) / ("commit", barrier) /
("use", what) operations; a round-robin scheduler with randomised order executes them against a model of mbarrier
phases.  Checked for random work lists: the schedule terminates (no deadlock), every raw stage is converted exactly
once before it is overwritten, every bf16 quad half is written before the MMA reads it and not overwritten before
the MMA's commit, accumulators are drained before reuse.

    python tools/experiments/pipeline_sim.py [trials]"""
import random
import sys

RAW, BF, SPLIT_WARPS = 3, 2, 8


class Bar:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0      # `phase` = number of completed phases

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, "too many arrivals in one phase"
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count

    def passed(self, parity):
        # mbarrier.try_wait.parity(p): true once the phase with parity p has completed, i.e. the barrier is now in a
        # phase of the opposite parity (a fresh barrier is in phase 0 -> waiting on parity 1 passes immediately)
        return (self.phase & 1) != parity


def simulate(items, seed, main_tf32=False):
    rng = random.Random(seed)
    full = [Bar(1) for _ in range(RAW)]
    empty = [Bar(SPLIT_WARPS + (1 if main_tf32 else 0)) for _ in range(RAW)]
    ready = [Bar(2 * SPLIT_WARPS) for _ in range(BF)]
    bf_empty = [Bar(1) for _ in range(BF)]
    tmem_full = [Bar(1) for _ in range(2)]
    tmem_empty = [Bar(4) for _ in range(2)]
    raw_content = [None] * RAW                  # (item, kb) currently held by a raw stage
    raw_reads = [0] * RAW
    raw_mma = [False] * RAW
    quad = [[None, None] for _ in range(BF)]     # per half: (item, kb) written, by how many warps
    quad_writes = [[0, 0] for _ in range(BF)]
    log = {"mma": [], "epi": []}

    def tma():
        it = 0
        for item, nkb in enumerate(items):
            for kb in range(nkb):
                s = it % RAW
                yield ("wait", empty[s], ((it // RAW) & 1) ^ 1)
                assert raw_content[s] is None or raw_reads[s] == SPLIT_WARPS, "raw stage overwritten before all splitter warps read it"
                assert raw_content[s] is None or not main_tf32 or raw_mma[s], "raw stage overwritten before the tf32 main MMAs read it"
                raw_content[s], raw_reads[s], raw_mma[s] = (item, kb), 0, False
                yield ("arrive", full[s])            # models expect_tx + the bytes landing
                it += 1

    def splitter(wid):
        it = jt = 0
        for item, nkb in enumerate(items):
            for kb in range(nkb):
                s, b, h = it % RAW, jt % BF, kb & 1
                if h == 0:
                    yield ("wait", bf_empty[b], ((jt // BF) & 1) ^ 1)
                yield ("wait", full[s], (it // RAW) & 1)
                assert raw_content[s] == (item, kb), f"splitter {wid} read raw stage {s} holding {raw_content[s]}, wanted {(item, kb)}"
                raw_reads[s] += 1
                if quad_writes[b][h] == 0 or quad[b][h] != (item, kb):
                    quad[b][h], quad_writes[b][h] = (item, kb), 0
                quad_writes[b][h] += 1
                last = kb == nkb - 1
                yield ("arrive", empty[s])
                yield ("arrive", ready[b])
                if last and h == 0:
                    yield ("arrive", ready[b])
                if h == 1 or last:
                    jt += 1
                it += 1

    def mma():
        it = jt = 0
        for item, nkb in enumerate(items):
            acc = item & 1
            yield ("wait", tmem_empty[acc], ((item >> 1) & 1) ^ 1)
            for kb in range(nkb):
                if main_tf32:
                    s = it % RAW
                    yield ("wait", full[s], (it // RAW) & 1)
                    assert raw_content[s] == (item, kb), f"tf32 main MMA read raw stage {s} holding {raw_content[s]}, wanted {(item, kb)}"
                    raw_mma[s] = True
                    yield ("arrive", empty[s])           # tcgen05.commit
                it += 1
                if (kb & 1) == 0 and kb != nkb - 1:
                    continue
                b, j = jt % BF, kb >> 1
                yield ("wait", ready[b], (jt // BF) & 1)
                halves = min(2, nkb - 2 * j)
                for h in range(halves):
                    assert quad[b][h] == (item, 2 * j + h) and quad_writes[b][h] == SPLIT_WARPS, \
                        f"MMA read quad {b} half {h}: {quad[b][h]} x{quad_writes[b][h]}, wanted {(item, 2 * j + h)}"
                log["mma"].append((item, j, halves))
                for h in range(2):
                    quad_writes[b][h] = 0
                yield ("arrive", bf_empty[b])            # tcgen05.commit
                jt += 1
            yield ("arrive", tmem_full[acc])

    def epilogue(w):
        for item, _ in enumerate(items):
            acc = item & 1
            yield ("wait", tmem_full[acc], (item >> 1) & 1)
            if w == 0:
                log["epi"].append(item)
            yield ("arrive", tmem_empty[acc])

    roles = [tma(), mma()] + [splitter(i) for i in range(SPLIT_WARPS)] + [epilogue(i) for i in range(4)]
    pending = [None] * len(roles)
    alive = set(range(len(roles)))
    idle_rounds = 0
    while alive:
        progressed = False
        order = list(alive)
        rng.shuffle(order)
        for r in order:
            op = pending[r]
            if op is None:
                try:
                    op = next(roles[r])
                except StopIteration:
                    alive.discard(r)
                    progressed = True
                    continue
            if op[0] == "wait":
                if op[1].passed(op[2]):
                    pending[r] = None
                    progressed = True
                else:
                    pending[r] = op
            else:
                op[1].arrive()
                pending[r] = None
                progressed = True
        idle_rounds = 0 if progressed else idle_rounds + 1
        assert idle_rounds < 3, f"deadlock with items={items}: waiting roles {[r for r in alive if pending[r]]}"
    assert log["epi"] == list(range(len(items)))
    assert [m[:2] for m in log["mma"]] == [(i, j) for i, n in enumerate(items) for j in range((n + 1) // 2)]


def main():
    trials = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    rng = random.Random(0)
    for t in range(trials):
        items = [rng.choice([1, 2, 3, 4, 5, 7, 8, 22, 23]) for _ in range(rng.randint(1, 6))]
        simulate(items, seed=t, main_tf32=False)
        simulate(items, seed=t, main_tf32=True)
    print(f"{trials} random work lists x 2 kernel variants: no deadlock, stage ownership respected")


if __name__ == "__main__":
    main()
