"""Host model of the slot protocol of the warp-specialised cooperative kernel (mcpar_b200/csrc/mh_coop.cuh).

The kernel's two roles -- mixture warps and owner warps -- walk the same iteration list and meet at named barriers
(FULL[s], DONE[s], EMPTY).  This test restates the two loops as generators over barrier operations, runs them under a
scheduler that only lets a `sync` proceed once the matching `arrive` has happened, and checks, for many
(batches, CTAs, steps per launch) shapes, that
  * neither role deadlocks and every barrier phase is balanced (one arrive, one sync),
  * every (batch, step) is proposed (P1), evaluated by the mixture warps and finished (P3) exactly once, in step order,
  * the shared partial buffer is never overwritten before the owners have read it (EMPTY), and a slot's proposal buffer
    never before the mixture warps are done with it (DONE).
The index arithmetic is the kernel's (cnt0 / cnt1, batch = cta + (2 bi + s) grid, the counters that replace m / nsteps)."""
import itertools

import pytest


def _counts(nbatch, grid, cta, nsteps):
    nb_cta = (nbatch - cta + grid - 1) // grid if cta < nbatch else 0
    return ((nb_cta + 1) >> 1) * nsteps, (nb_cta >> 1) * nsteps


def mixture(nbatch, grid, cta, nsteps, log):
    cnt0, cnt1 = _counts(nbatch, grid, cta, nsteps)
    seq = 0
    for i in range(2 * cnt0):
        s, m = i & 1, i >> 1
        if m >= (cnt1 if s else cnt0):
            continue
        yield ("sync", "FULL%d" % s)
        log.append(("P2", s, m))                    # reads sx[s], computes the partials in registers
        if seq > 0:
            yield ("sync", "EMPTY")
        log.append(("store_sq", s, m))
        yield ("arrive", "DONE%d" % s)
        seq += 1


def owners(nbatch, grid, cta, nsteps, log):
    cnt0, cnt1 = _counts(nbatch, grid, cta, nsteps)
    total_active = cnt0 + cnt1
    seq = 0
    state = {0: [0, -1], 1: [0, -1]}                # per slot: (batch index within the slot, step) of the iteration being finished
    for i in range(-2, 2 * cnt0):
        s, m = i & 1, i >> 1
        cnt = cnt1 if s else cnt0
        do_p3, do_p1 = i >= 0 and m < cnt, m + 1 < cnt
        if not do_p3 and not do_p1:
            continue
        if do_p3:
            bi, k = state[s]
            assert (bi, k) == (m // nsteps, m % nsteps)     # the counters replace the divisions
            yield ("sync", "DONE%d" % s)
            log.append(("read_sq", s, m))
            if seq < total_active - 1:
                yield ("arrive", "EMPTY")
            seq += 1
            log.append(("P3", cta + (2 * bi + s) * grid, k))
        if do_p1:
            b1, k1 = state[s][0], state[s][1] + 1
            if k1 == nsteps:
                k1, b1 = 0, b1 + 1
            state[s] = [b1, k1]
            log.append(("P1", cta + (2 * b1 + s) * grid, k1, s))
            yield ("arrive", "FULL%d" % s)


def run_cta(nbatch, grid, cta, nsteps, owners_first=False):
    log = []
    roles = {"mix": mixture(nbatch, grid, cta, nsteps, log), "own": owners(nbatch, grid, cta, nsteps, log)}
    pending = {}                                    # role -> barrier it waits at
    arrived = {}                                    # barrier -> outstanding arrivals
    live = set(roles)
    while live:
        progressed = False
        for name in sorted(live, reverse=owners_first):   # either role may run ahead as far as its barriers let it
            if name in pending:
                b = pending[name]
                if arrived.get(b, 0) > 0:
                    arrived[b] -= 1
                    del pending[name]
                else:
                    continue
            try:
                while True:
                    op, b = next(roles[name])
                    progressed = True
                    if op == "arrive":
                        arrived[b] = arrived.get(b, 0) + 1
                        assert arrived[b] == 1, "barrier %s armed twice before its sync" % b
                    elif arrived.get(b, 0) > 0:
                        arrived[b] -= 1
                    else:
                        pending[name] = b
                        break
            except StopIteration:
                live.discard(name)
                progressed = True
        assert progressed, "deadlock: %s" % pending
    assert not any(arrived.values()), "unbalanced barrier phases: %s" % arrived
    return log


@pytest.mark.parametrize("owners_first", [False, True])
@pytest.mark.parametrize("nbatch,grid,nsteps", list(itertools.product([1, 2, 3, 7, 8, 37], [1, 2, 5], [1, 3, 10])))
def test_slot_protocol_covers_every_batch_step_once(nbatch, grid, nsteps, owners_first):
    done = []
    for cta in range(min(grid, nbatch)):
        log = run_cta(nbatch, grid, cta, nsteps, owners_first)
        p1 = [(e[1], e[2]) for e in log if e[0] == "P1"]
        p3 = [(e[1], e[2]) for e in log if e[0] == "P3"]
        assert sorted(p1) == sorted(p3) and len(set(p3)) == len(p3)
        for b in {b for b, _ in p3}:                # steps of a batch are finished in order, each proposed before it is finished
            ks = [k for bb, k in p3 if bb == b]
            assert ks == list(range(nsteps))
        pos1 = {e[1:3]: n for n, e in enumerate(log) if e[0] == "P1"}
        pos3 = {e[1:3]: n for n, e in enumerate(log) if e[0] == "P3"}
        assert all(pos1[key] < pos3[key] for key in pos3)
        # the shared partial buffer: every store is read before the next store
        seq = [e[0] for e in log if e[0] in ("store_sq", "read_sq")]
        assert seq == ["store_sq", "read_sq"] * (len(seq) // 2)
        # a slot's proposal buffer: P1 into slot s only after the mixture warps have consumed the previous content (P2)
        for s in (0, 1):
            ev = [e[0] for e in log if (e[0] == "P2" and e[1] == s) or (e[0] == "P1" and e[3] == s)]
            assert ev == ["P1", "P2"] * (len(ev) // 2)
        done += [b for b, k in p3 if k == nsteps - 1]
    assert sorted(done) == list(range(nbatch))
