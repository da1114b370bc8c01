"""Exhaustive model check of the flag protocol of the fused halo (molchanica_b200/csrc/halo_sync.cuh, integrate.cu,
pair_force.cu): every interleaving of the ranks' kernel phases, for rings of 2 and 3 ranks over several steps with a
rebuild in between, must (1) never deadlock, (2) let a pair kernel read exactly the ghosts of its own epoch, and (3)
never let a push overwrite ghosts the neighbour's pair kernel of the previous epoch has not read yet.

Model of one rank's stream for epoch e (phases run in order, a wait blocks until its condition holds):
  push step:    K1 ack(e-1) -> both neighbours      K2 wait ack >= e-1 from both     K3 store ghosts(e) into both
                K4 ready(e) -> both neighbours      P1 wait ready >= e from both     P2 read ghosts, check version
  rebuild step: K1 ack(e-1)                         R  collective (all ranks), ghosts := e          P2 read ghosts
(K2 only gates the blocks that push, P1 only the blocks that read ghosts; modelling them as one phase each is the
conservative choice for deadlock and exactly right for the data checks.)"""
import itertools

import pytest


def phases(schedule):
    seq = []
    for e, rebuild in schedule:
        if rebuild:
            seq += [("K1", e), ("R", e), ("P2", e)]
        else:
            seq += [("K1", e), ("K2", e), ("K3", e), ("K4", e), ("P1", e), ("P2", e)]
    return seq


def explore(n, schedule):
    seq = phases(schedule)
    prev = [(r - 1) % n for r in range(n)]
    nxt = [(r + 1) % n for r in range(n)]
    # state: pc per rank; ack_from_prev/next, ready_from_prev/next per rank; ghost version from prev / next per rank;
    # last epoch whose ghosts the rank has read
    init = (tuple([0] * n), tuple([1] * n), tuple([1] * n), tuple([0] * n), tuple([0] * n), tuple([1] * n), tuple([1] * n),
            tuple([1] * n))
    seen, stack, finals = {init}, [init], 0
    while stack:
        pc, ackp, ackn, rdyp, rdyn, gvp, gvn, read = stack.pop()
        if all(p == len(seq) for p in pc):
            finals += 1
            continue
        progressed = False
        at_r = [p < len(seq) and seq[p][0] == "R" for p in pc]
        for r in range(n):
            if pc[r] == len(seq):
                continue
            kind, e = seq[pc[r]]
            ackp_, ackn_, rdyp_, rdyn_, gvp_, gvn_, read_ = map(list, (ackp, ackn, rdyp, rdyn, gvp, gvn, read))
            if kind == "K1":      # towards prev this rank is "next", towards next it is "prev"
                ackn_[prev[r]] = max(ackn_[prev[r]], e - 1)
                ackp_[nxt[r]] = max(ackp_[nxt[r]], e - 1)
            elif kind == "K2":
                if ackp[r] < e - 1 or ackn[r] < e - 1:
                    continue
            elif kind == "K3":
                # the neighbours must have read the ghosts of the previous epoch before they are overwritten
                assert read[prev[r]] >= e - 1 and read[nxt[r]] >= e - 1, ("overwrite before read", r, e, read)
                gvn_[prev[r]] = e      # my first layer is prev's ghost-next block
                gvp_[nxt[r]] = e       # my last layer is next's ghost-prev block
            elif kind == "K4":
                rdyn_[prev[r]] = max(rdyn_[prev[r]], e)
                rdyp_[nxt[r]] = max(rdyp_[nxt[r]], e)
            elif kind == "P1":
                if rdyp[r] < e or rdyn[r] < e:
                    continue
            elif kind == "P2":
                assert gvp[r] == e and gvn[r] == e, ("stale or future ghosts", r, e, gvp[r], gvn[r])
                read_[r] = e
            elif kind == "R":
                if not all(at_r):
                    continue           # collective: proceeds only when every rank has arrived
                if r != 0:
                    continue           # fire it once, for all ranks together
                pc2 = tuple(p + 1 for p in pc)
                st = (pc2, ackp, ackn, rdyp, rdyn, tuple([e] * n), tuple([e] * n), read)
                progressed = True
                if st not in seen:
                    seen.add(st)
                    stack.append(st)
                continue
            pc2 = tuple(p + 1 if i == r else p for i, p in enumerate(pc))
            st = (pc2, tuple(ackp_), tuple(ackn_), tuple(rdyp_), tuple(rdyn_), tuple(gvp_), tuple(gvn_), tuple(read_))
            progressed = True
            if st not in seen:
                seen.add(st)
                stack.append(st)
        assert progressed, ("deadlock", pc, [seq[p] if p < len(seq) else None for p in pc])
    return len(seen), finals


@pytest.mark.parametrize("n", [2, 3])
def test_every_interleaving_is_safe_and_live(n):
    # epochs start at 2 (the engine starts its counter at 1 so that the first acks are real signals); step 4 rebuilds
    schedule = [(2, False), (3, False), (4, True), (5, False), (6, False)]
    states, finals = explore(n, schedule)
    assert finals >= 1 and states > 100


def _explore_without(kind, n, schedule):
    global phases
    orig = phases
    phases = lambda sch: [p for p in orig(sch) if p[0] != kind]
    try:
        return explore(n, schedule)
    finally:
        phases = orig


@pytest.mark.parametrize("kind,what", [("K2", "overwrite before read"), ("P1", "stale or future ghosts")])
def test_the_model_catches_a_missing_wait(kind, what):
    """Sanity of the checker itself: without the ack wait (K2) a fast rank overwrites ghosts its neighbour has not read;
    without the ready wait (P1) a pair kernel reads ghosts of the wrong epoch."""
    with pytest.raises(AssertionError) as err:
        _explore_without(kind, 2, [(2, False), (3, False), (4, False)])
    assert what in str(err.value)
