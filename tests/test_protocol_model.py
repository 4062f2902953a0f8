"""A small discrete-event model of the release protocol of attn_bwd_pp_kernel (finetrainers_b200/csrc/b2d_attn.cu): three
consumer warpgroups of four warps, four S/dP buffers, one MMA-issuing thread.  Tile `it` belongs to warpgroup it % 3 and
buffer it % 4; a warpgroup releases a tile on ONE mbarrier that just counts arrivals (4 warps = 128 threads) and is shared
by tiles it, it + 3, ...; the MMA thread issues S/dP(it + 4) when tile `it` is released, i.e. S/dP(it + 3) is issued when
tile it - 1 - another warpgroup's - is released.

The model checks the invariant the accumulation GEMM relies on - when the release phase of tile `it` completes, all four warps
of its warpgroup have finished writing tile `it` - under random and adversarial warp delays, with and without the
warpgroup-wide barrier at the start of every tile.  Without it a warp that runs ahead arrives twice before a delayed sibling
arrives once (the bug of profiles/r2b_attention_backward_nan.md); with it the invariant holds and nothing deadlocks."""
import heapq
import random

NWG, NWARP, NBUF = 3, 4, 4


def simulate(n_tiles, seed, wg_barrier, stall=None):
    """-> (violations, finished).  stall = (wg, warp, tile, extra_time): one adversarial delay."""
    rng = random.Random(seed)
    t_mma = 1.0                                   # time the tensor pipe needs per issued group
    s_ready = {}                                  # tile -> time its S/dP is readable
    arrivals = [0] * NWG                          # arrival count of ds_full[wg] in its current phase
    phase_done = [[] for _ in range(NWG)]         # completion times of the phases of ds_full[wg]
    written = {}                                  # (tile, warp) -> time the warp finished writing the tile
    violations = []
    # events: (time, seq, kind, payload)
    ev, seq = [], 0

    def push(t, kind, payload):
        nonlocal seq
        heapq.heappush(ev, (t, seq, kind, payload))
        seq += 1

    pipe_free = 0.0

    def issue_sdp(now, tile):
        nonlocal pipe_free
        pipe_free = max(pipe_free, now) + t_mma
        s_ready[tile] = pipe_free

    for tile in range(min(NBUF, n_tiles)):
        issue_sdp(0.0, tile)
    mma_next = 0                                  # next tile whose release the MMA thread waits for
    # per-warp progress: index into its tile list; per-warpgroup barrier generation bookkeeping
    tiles_of = {wg: list(range(wg, n_tiles, NWG)) for wg in range(NWG)}
    pos = {(wg, w): 0 for wg in range(NWG) for w in range(NWARP)}
    at_barrier = {wg: {} for wg in range(NWG)}   # tile -> set of warps waiting at the start-of-tile barrier
    for wg in range(NWG):
        for w in range(NWARP):
            push(0.0, "start_tile", (wg, w))
    finished_warps = 0
    mma_done = False
    guard = 0
    while ev:
        guard += 1
        assert guard < 2_000_000
        now, _, kind, payload = heapq.heappop(ev)
        if kind == "start_tile":
            wg, w = payload
            if pos[(wg, w)] >= len(tiles_of[wg]):
                finished_warps += 1
                continue
            tile = tiles_of[wg][pos[(wg, w)]]
            if wg_barrier:
                waiting = at_barrier[wg].setdefault(tile, set())
                waiting.add(w)
                if len(waiting) < NWARP:
                    continue                      # parked until the last sibling shows up
                for w2 in waiting:
                    push(now, "wait_s", (wg, w2, tile))
            else:
                push(now, "wait_s", (wg, w, tile))
        elif kind == "wait_s":
            wg, w, tile = payload
            if tile not in s_ready:               # S/dP not issued yet: poll again when something changes
                push(now + 0.25, "wait_s", payload)
                continue
            t0 = max(now, s_ready[tile])
            work = 1.5 + rng.random()             # consume the tile
            if stall and stall[:3] == (wg, w, tile):
                work += stall[3]
            push(t0 + work, "arrive", (wg, w, tile))
        elif kind == "arrive":
            wg, w, tile = payload
            written[(tile, w)] = now
            arrivals[wg] += 1
            if arrivals[wg] == NWARP:             # the phase completes on the 4th arrival, whoever made it
                arrivals[wg] = 0
                phase_done[wg].append(now)
            pos[(wg, w)] += 1
            push(now, "start_tile", (wg, w))
            push(now, "mma_poll", None)
        elif kind == "mma_poll":
            # the MMA thread handles releases strictly in tile order
            while mma_next < n_tiles:
                wg, k = mma_next % NWG, mma_next // NWG
                if len(phase_done[wg]) <= k:
                    break
                # release of tile mma_next observed: every warp of the warpgroup must have written it
                missing = [w for w in range(NWARP) if (mma_next, w) not in written or written[(mma_next, w)] > phase_done[wg][k]]
                if missing:
                    violations.append((mma_next, missing))
                t = max(now, phase_done[wg][k])
                pipe_free = max(pipe_free, t) + t_mma          # accumulation GEMMs of the tile
                if mma_next + NBUF < n_tiles:
                    issue_sdp(pipe_free, mma_next + NBUF)
                mma_next += 1
            if mma_next == n_tiles:
                mma_done = True
    return violations, (mma_done and finished_warps == NWG * NWARP)


def test_release_protocol_needs_the_warpgroup_barrier():
    # a warp delayed by more than a tile period: without the barrier a sibling laps it and the phase completes without it
    stall = (1, 2, 4, 12.0)      # warpgroup 1, warp 2, its tile 4
    v, done = simulate(40, seed=0, wg_barrier=False, stall=stall)
    assert v and any(2 in missing for _, missing in v), "the model must reproduce the double arrival"
    # with the barrier: same adversarial delay, no violation, no deadlock
    v, done = simulate(40, seed=0, wg_barrier=True, stall=stall)
    assert not v and done


def test_release_protocol_with_barrier_holds_under_random_delays():
    for seed in range(200):
        rng = random.Random(1000 + seed)
        stall = (rng.randrange(NWG), rng.randrange(NWARP), None, rng.choice([0.0, 3.0, 8.0, 20.0]))
        n = rng.choice([5, 12, 56])
        tile = rng.choice(list(range(stall[0], n, NWG)))
        v, done = simulate(n, seed=seed, wg_barrier=True, stall=(stall[0], stall[1], tile, stall[3]))
        assert not v and done, (seed, n, stall, v[:3])
