"""Executable model of the synchronisation protocol of the fused NVLink all-reduce (csrc/tp_allreduce.cu,
engine/peer_reduce.py), run with one Python thread per rank and adversarial delays.

What the CUDA kernel relies on, restated here step for step:
  * every rank owns TWO data buffers that alternate between consecutive exchanges, and one flag word per peer;
  * exchange number e on rank r:  (1) the GEMM writes r's partial into buffer e % 2      [previous kernel in the stream]
                                  (2) r stores e + 1 into flags_of(p)[r] for every peer p  [st.release.sys, one thread per
                                      peer, all at once: the order of these stores is immaterial to the protocol]
                                  (3) r waits until flags_of(r)[p] >= e + 1 for every p    [ld.acquire.sys]
                                  (4) r reads buffer e % 2 of every rank and reduces
  * nothing else orders the ranks: a fast rank may run ahead as far as the flags let it.
The property to hold: in step (4) of exchange e every rank reads exactly the partials of exchange e -- no rank has
overwritten a buffer a slower peer is still reading.  The model checks it with payloads that carry (rank, exchange),
and a negative control (ONE buffer instead of two) shows the check has teeth: it must catch the overwrite.
"""
import random
import threading
import time

import pytest


class Fabric:
    def __init__(self, world: int, buffers: int):
        self.world, self.nbuf = world, buffers
        self.data = [[None] * buffers for _ in range(world)]          # data[rank][buf] = (rank, exchange)
        self.flags = [[0] * world for _ in range(world)]              # flags[owner][writer]
        self.errors: list[str] = []
        self.deadline = time.monotonic() + 20.0


def rank_main(fab: Fabric, rank: int, exchanges: int, delay):
    for e in range(exchanges):
        buf = e % fab.nbuf
        delay(rank, e, "before_gemm")
        fab.data[rank][buf] = (rank, e)                               # (1) this rank's partial
        for p in range(fab.world):                                    # (2) announce
            if p != rank:
                fab.flags[p][rank] = e + 1
        for p in range(fab.world):                                    # (3) wait for every peer's announcement
            while p != rank and fab.flags[rank][p] < e + 1:
                if time.monotonic() > fab.deadline:
                    fab.errors.append(f"rank {rank} stuck in exchange {e} waiting for {p}")
                    return
                time.sleep(0)
        delay(rank, e, "before_read")
        for p in range(fab.world):                                    # (4) read all partials of this exchange
            got = fab.data[p][buf]
            if got != (p, e):
                fab.errors.append(f"rank {rank}, exchange {e}: buffer of rank {p} holds {got}")


def run(world, buffers, exchanges, delay):
    fab = Fabric(world, buffers)
    threads = [threading.Thread(target=rank_main, args=(fab, r, exchanges, delay)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(30)
    assert not any(t.is_alive() for t in threads), "model deadlocked"
    return fab.errors


@pytest.mark.parametrize("world", [2, 4, 8])
def test_two_buffers_survive_adversarial_skew(world):
    rnd = random.Random(world)
    lock = threading.Lock()

    def jitter(rank, e, where):
        with lock:
            r = rnd.random()
        if r < 0.2:
            time.sleep(0.002 * r * 10)

    assert run(world, 2, 120, jitter) == []

    def slow_reader(rank, e, where):                                   # one rank always reads late, the others race ahead
        if rank == 0 and where == "before_read":
            time.sleep(0.003)

    assert run(world, 2, 60, slow_reader) == []

    def slow_writer(rank, e, where):                                   # one rank always arrives late
        if rank == world - 1 and where == "before_gemm":
            time.sleep(0.003)

    assert run(world, 2, 60, slow_writer) == []


def test_single_buffer_is_caught():
    """Negative control: with one buffer a fast rank's next GEMM overwrites what a slow reader has not read yet."""
    def slow_reader(rank, e, where):
        if rank == 0 and where == "before_read":
            time.sleep(0.01)

    errors = run(2, 1, 20, slow_reader)
    assert errors and "holds" in errors[0]
