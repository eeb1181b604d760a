"""Model check of the neighbour handshakes of the library-owned halo exchange (csrc/halo_device.cuh, csrc/k_halo.cu):
the pull protocol first, the push protocol of the mixed exchange below.

The kernels cannot be run on several GPUs in this container, but their PROTOCOL can be executed: every rank is a
sequence of micro-operations, a random scheduler interleaves the ranks, and the invariants a ping-pong time loop
needs are asserted on every read and write:

  step n of rank r:   announce   flags[p][r] = n for every neighbour p        (st.release.sys, handshake kernel / block 0)
                      pull       for every link: wait flags[r][p] >= n, then read p's field      (ld.acquire.sys)
                      compute    write field version n into the OTHER buffer                      (the stencil)

  * a pull at step n must see exactly version n-1 of the neighbour's field, in the buffer that holds it;
  * nobody may be overwriting a buffer while a neighbour still has to read it;
  * no deadlock: the scheduler always finds a runnable rank until every rank has finished.

The same model with the wait removed must FAIL, or the check would prove nothing."""
import random

import pytest

from b200stencil.halo.partitioner import CubedSpherePartitioner, layout_for


def gpu_neighbours(n_gpus):
    part = CubedSpherePartitioner(24, layout_for(n_gpus), 3)
    nb = {g: set() for g in range(n_gpus)}
    for l in part.all_links():
        a, b = part.gpu_of(l.src, n_gpus), part.gpu_of(l.dst, n_gpus)
        if a != b:
            nb[b].add(a)  # b pulls from a
    return nb


class Violation(AssertionError):
    pass


def run_schedule(nb, steps, rng, wait=True):
    world = len(nb)
    flags = [[0] * world for _ in range(world)]           # flags[r][p]: latest epoch announced by p to r
    version = [[0, -1] for _ in range(world)]             # version[r][buf]: field version held by each ping-pong buffer
    writing = [None] * world                              # buffer a rank is overwriting right now
    # per-rank program counter: (step, phase, pending links)
    pc = [{"n": 1, "phase": "announce", "todo": None} for _ in range(world)]
    done = 0
    guard = 0
    while done < world:
        guard += 1
        assert guard < 200000, "scheduler did not terminate"
        runnable = []
        for r in range(world):
            st = pc[r]
            if st["n"] > steps:
                continue
            if st["phase"] == "pull" and wait:
                if any(flags[r][p] >= st["n"] for p in st["todo"]):
                    runnable.append(r)
            else:
                runnable.append(r)
        if not runnable:
            raise Violation("deadlock: every unfinished rank is waiting")
        r = rng.choice(runnable)
        st = pc[r]
        n = st["n"]
        if st["phase"] == "announce":
            for p in nb[r]:
                flags[p][r] = n
            st["phase"], st["todo"] = "pull", sorted(nb[r])
        elif st["phase"] == "pull":
            ready = [p for p in st["todo"] if flags[r][p] >= n] if wait else list(st["todo"])
            p = rng.choice(ready)
            src_buf = (n - 1) % 2                        # version n-1 lives in buffer (n-1) % 2
            if writing[p] == src_buf:
                raise Violation(f"rank {r} step {n}: reads buffer {src_buf} of rank {p} while {p} overwrites it")
            if version[p][src_buf] != n - 1:
                raise Violation(f"rank {r} step {n}: expected version {n - 1} of rank {p}, found {version[p][src_buf]}")
            st["todo"].remove(p)
            if not st["todo"]:
                st["phase"] = "compute_begin"
        elif st["phase"] == "compute_begin":
            writing[r] = n % 2
            version[r][n % 2] = None                     # being overwritten: neither the old nor the new version
            st["phase"] = "compute_end"
        else:  # compute_end
            version[r][n % 2] = n
            writing[r] = None
            st["n"], st["phase"] = n + 1, "announce"
            if st["n"] > steps:
                done += 1
    return True


@pytest.mark.parametrize("n_gpus", [2, 4, 8])
def test_handshake_orders_pulls_and_overwrites(n_gpus):
    nb = gpu_neighbours(n_gpus)
    for g, s in nb.items():  # the protocol needs symmetric adjacency: whoever I wait for also waits for me
        assert s and all(g in nb[p] for p in s)
    rng = random.Random(20240724 + n_gpus)
    for _ in range(300):
        assert run_schedule(nb, steps=5, rng=rng)


def test_without_the_wait_the_model_catches_the_race():
    nb = gpu_neighbours(4)
    rng = random.Random(1)
    caught = 0
    for _ in range(200):
        try:
            run_schedule(nb, steps=5, rng=rng, wait=False)
        except Violation:
            caught += 1
    assert caught > 150


# ---- the push protocol of the mixed exchange (csrc/halo_device.cuh, version 3) ---------------------------------------
#
#   step n of rank r:   announce   A[p][r] = n for every neighbour p: "my field of version n-1 is final AND the halos
#                                  of the buffer that holds it may be overwritten"
#                       push       for every neighbour p: wait A[r][p] >= n, write my interior (version n-1) into p's
#                                  halo of buffer (n-1) % 2, then D[p][r] = n: "delivered"
#                       wait       D[r][p] >= n for every neighbour p
#                       compute    read buffer (n-1) % 2 (interior + halos), write version n into the OTHER buffer
#
#   * a push must not land in a halo its owner is still reading (the previous user of that buffer is compute n-2);
#   * compute n must find version n-1 in every halo of the buffer it reads;
#   * no deadlock.  With either wait removed the model must FAIL.


def run_push_schedule(nb, steps, rng, wait_a=True, wait_d=True, buffers=2):
    world = len(nb)
    A = [[0] * world for _ in range(world)]
    D = [[0] * world for _ in range(world)]
    version = [[0, -1] for _ in range(world)]                       # interior version per ping-pong buffer
    halo = [[{p: None for p in nb[r]} for _ in range(2)] for r in range(world)]  # halo[r][buf][p]: version pushed by p
    single = buffers == 1  # the stencil reads the SAME field every step and writes elsewhere (bench.py's step)
    reading = [None] * world                                       # buffer a rank's stencil is reading right now
    pc = [{"n": 1, "phase": "announce", "todo": None} for _ in range(world)]
    done, guard = 0, 0
    while done < world:
        guard += 1
        assert guard < 400000, "scheduler did not terminate"
        runnable = []
        for r in range(world):
            st = pc[r]
            if st["n"] > steps:
                continue
            if st["phase"] == "push" and wait_a:
                if any(A[r][p] >= st["n"] for p in st["todo"]):
                    runnable.append(r)
            elif st["phase"] == "wait" and wait_d:
                if all(D[r][p] >= st["n"] for p in nb[r]):
                    runnable.append(r)
            else:
                runnable.append(r)
        if not runnable:
            raise Violation("deadlock: every unfinished rank is waiting")
        r = rng.choice(runnable)
        st = pc[r]
        n = st["n"]
        buf = 0 if single else (n - 1) % 2
        if st["phase"] == "announce":
            for p in nb[r]:
                A[p][r] = n
            st["phase"], st["todo"] = "push", sorted(nb[r])
        elif st["phase"] == "push":
            ready = [p for p in st["todo"] if A[r][p] >= n] if wait_a else list(st["todo"])
            p = rng.choice(ready)
            if not single and version[r][buf] != n - 1:
                raise Violation(f"rank {r} step {n}: pushes version {version[r][buf]} instead of {n - 1}")
            if reading[p] == buf:
                raise Violation(f"rank {r} step {n}: writes into halo buffer {buf} of rank {p} while {p}'s stencil reads it")
            halo[p][buf][r] = n - 1
            D[p][r] = n
            st["todo"].remove(p)
            if not st["todo"]:
                st["phase"] = "wait"
        elif st["phase"] == "wait":
            st["phase"] = "compute_begin"
        elif st["phase"] == "compute_begin":
            for p in nb[r]:
                if halo[r][buf][p] != n - 1:
                    raise Violation(f"rank {r} step {n}: halo from rank {p} holds version {halo[r][buf][p]}, not {n - 1}")
            reading[r] = buf
            if not single:
                version[r][n % 2] = None
            st["phase"] = "compute_end"
        else:  # compute_end
            if not single:
                version[r][n % 2] = n
            reading[r] = None
            st["n"], st["phase"] = n + 1, "announce"
            if st["n"] > steps:
                done += 1
    return True


@pytest.mark.parametrize("buffers", [1, 2])
@pytest.mark.parametrize("n_gpus", [2, 4, 8])
def test_push_protocol_orders_deliveries_and_reads(n_gpus, buffers):
    nb = gpu_neighbours(n_gpus)
    rng = random.Random(20241018 + n_gpus)
    for _ in range(300):
        assert run_push_schedule(nb, steps=5, rng=rng, buffers=buffers)


@pytest.mark.parametrize("drop", ["announcement", "delivery"])
def test_push_protocol_model_catches_a_missing_wait(drop):
    """Deliveries alone keep a rank within one step of its neighbours, which is enough for a two-buffer time loop (the
    push lands in the buffer the neighbour is NOT reading); with ONE buffer read every step (bench.py) the announcement
    is what keeps a push out of a halo that is still being read."""
    nb = gpu_neighbours(4)
    rng = random.Random(7)
    caught = 0
    for _ in range(200):
        try:
            run_push_schedule(nb, steps=5, rng=rng, wait_a=drop != "announcement", wait_d=drop != "delivery", buffers=1)
        except Violation:
            caught += 1
    assert caught > 150, caught
