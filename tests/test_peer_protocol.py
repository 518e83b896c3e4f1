"""Model check (CPU, no CUDA) of the mailbox protocol of the peer-memory exchange, primitive3d_b200/csrc/mc_peer.cu:

  call e of rank r (parity p = e & 1):
    export   for every rank t: data[t][p][r] = payload(e, r)      (stores into the other ranks' mailboxes)
             for every rank t: flags[t][p][r] = e                   (after a fence: a flag never overtakes its payload)
    wait     until flags[r][p][t] == e for every t
    read     data[r][p][t] must be payload(e, t) for every t         (k_apply_exchange, face pass, the host's readback)

The claim in the kernel's header: TWO parities make a mailbox safe against a neighbour that runs ahead, because a rank
can only deliver call e + 2 after its wait of call e + 1, i.e. after every other rank's export of call e + 1, which that
rank's stream runs after everything that reads call e.  The model runs the ranks as interleaved step machines under
random and adversarial schedules and checks every read; with ONE buffer the same schedules do corrupt reads, so the
check has teeth."""
import random

import pytest


def rank_program(r, world, calls, buffers, data, flags, log):
    """Generator: one step per memory operation of rank r; yields False while spinning."""
    for e in range(1, calls + 1):
        p = e % buffers
        for t in range(world):                       # export: payload stores
            data[t][p][r] = (e, r)
            yield True
        for t in range(world):                       # export: flags (release order: after all payload stores)
            flags[t][p][r] = e
            yield True
        for t in range(world):                       # wait
            while flags[r][p][t] != e:
                # with one buffer a flag can jump past e: the real kernel would spin into its timeout; count it
                if flags[r][p][t] > e:
                    log.append(("overrun", r, e, t))
                    break
                yield False
            yield True
        for t in range(world):                       # read
            if data[r][p][t] != (e, t):
                log.append(("corrupt", r, e, t, data[r][p][t]))
            yield True


def run(world, calls, buffers, pick):
    data = [[[None] * world for _ in range(buffers)] for _ in range(world)]
    flags = [[[0] * world for _ in range(buffers)] for _ in range(world)]
    log = []
    progs = {r: rank_program(r, world, calls, buffers, data, flags, log) for r in range(world)}
    spinning = set()   # ranks whose last step was a failed flag test
    steps = 0
    while progs:
        r = pick(sorted(progs), spinning)
        try:
            moved = next(progs[r])
            (spinning.discard if moved else spinning.add)(r)
        except StopIteration:
            del progs[r]
            spinning.discard(r)
        steps += 1
        assert steps < 2_000_000, "the model does not terminate: a rank waits for ever"
    return log


@pytest.mark.parametrize("world", [2, 3, 8])
def test_two_buffers_are_enough_under_random_schedules(world):
    for seed in range(60):
        rng = random.Random(seed)
        # a biased scheduler: one rank gets most of the steps, so it runs ahead as far as the protocol lets it
        fast = rng.randrange(world)
        pick = lambda alive, spinning: fast if fast in alive and rng.random() < 0.8 else rng.choice(alive)
        assert run(world, calls=6, buffers=2, pick=pick) == []


def test_two_buffers_under_a_run_ahead_schedule():
    # always advance the highest rank that is not spinning (it runs as far ahead as the protocol lets it), then the
    # lowest: the two extremes of "who is ahead of whom" at a shard boundary
    for world in (2, 4, 8):
        for order in (-1, 0):
            def pick(alive, spinning):
                free = [r for r in alive if r not in spinning]
                if not free:           # everybody failed its last flag test: test them all again
                    spinning.clear()
                    free = alive
                return free[order]
            assert run(world, calls=5, buffers=2, pick=pick) == []


def test_one_buffer_is_not_enough():
    bad = 0
    for seed in range(60):
        rng = random.Random(seed)
        fast = rng.randrange(3)
        pick = lambda alive, spinning: fast if fast in alive and rng.random() < 0.8 else rng.choice(alive)
        bad += bool(run(3, calls=6, buffers=1, pick=pick))
    assert bad > 0, "a single buffer survived every schedule: the model does not exercise the hazard"
