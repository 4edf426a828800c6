"""CPU property test of the scoring kernel's selection LOGIC (gat-recommendation_b200/csrc/score_tc.cu), restated in
numpy: per row and column half a list of the K largest 16-column piece maxima; a piece is dumped when its maximum
beats its own half's threshold ('>') and reaches the other half's ('>='); thresholds may be arbitrarily stale
(lockstep merges, the other warp's value read through shared memory) and the two halves may run at any relative
speed; the select step keeps the dumped scores >= tau = max(final thresholds) and takes the top-k by (score desc,
id asc).  Whatever the interleaving and staleness, the result must be the exact top-k of the row with ties to the
lower id (etpgt/model/base.py:59-78 through torch.topk's documented order on distinct values; the tie rule is
BASELINE.json's).  The GPU tests check the kernel against the oracle; this one checks that the rule itself cannot
skip a top-k item, over thousands of adversarial schedules, without a GPU."""

import numpy as np

PIECE, TILE = 16, 256


def exact_topk(scores, k):
    order = np.lexsort((np.arange(scores.size), -scores))      # score descending, id ascending
    return order[:k]


def emulate(scores, k, rng, merge_at=8):
    n = scores.size
    tiles = (n + TILE - 1) // TILE
    padded = np.full(tiles * TILE, -np.inf)
    padded[:n] = scores
    pieces = padded.reshape(tiles, 2, TILE // 2 // PIECE, PIECE)          # [tile, half, piece, column]
    best = [[], []]                    # the K largest piece maxima merged so far, per half
    pending = [[], []]
    thr = [-np.inf, -np.inf]           # each half's OWN threshold as of its last merge
    published = [[-np.inf], [-np.inf]]  # history of published thresholds (the other half may read a stale one)
    pos = [0, 0]                       # next (tile * pieces_per_half + piece) of each half
    per_half = TILE // 2 // PIECE
    dumped = []

    def merge(h):
        merged = sorted(best[h] + pending[h], reverse=True)[:k]
        best[h], pending[h] = merged, []
        if len(merged) == k:
            thr[h] = merged[-1]
            published[h].append(thr[h])

    while pos[0] < tiles * per_half or pos[1] < tiles * per_half:
        live = [h for h in (0, 1) if pos[h] < tiles * per_half]
        # any relative speed — but a half is never more than two tiles ahead (two accumulator stages)
        h = int(rng.choice(live))
        other = 1 - h
        if pos[h] // per_half > pos[other] // per_half + 1 and other in live:
            h, other = other, h
        t, p = divmod(pos[h], per_half)
        row = pieces[t, h, p]
        mx = row.max()
        # the other half's threshold: any value it has published so far (possibly stale)
        seen = published[other][int(rng.integers(0, len(published[other])))]
        if mx > thr[h] and mx >= seen and mx > -np.inf:
            col0 = t * TILE + h * (TILE // 2) + p * PIECE
            dumped.append((col0, row.copy()))
            pending[h].append(mx)
        pos[h] += 1
        if pos[h] % per_half == 0 and (len(pending[h]) >= merge_at or rng.random() < 0.1):   # one merge check per tile
            merge(h)
    merge(0)
    merge(1)
    tau = max(thr)
    cand = [(v, col0 + j) for col0, row in dumped for j, v in enumerate(row) if v >= tau and v > -np.inf]
    cand.sort(key=lambda c: (-c[0], c[1]))
    return np.array([c[1] for c in cand[:k]]), len(dumped)


def test_piece_dump_rule_is_exact_for_any_schedule():
    rng = np.random.default_rng(0)
    cases = 0
    for trial in range(400):
        n = int(rng.integers(300, 6000))
        kind = trial % 4
        if kind == 0:
            scores = rng.standard_normal(n)
        elif kind == 1:                                     # heavy ties: small integers
            scores = rng.integers(-3, 4, size=n).astype(np.float64)
        elif kind == 2:                                     # adversarial order: slowly increasing with ties
            scores = np.floor(np.arange(n) / 7.0) + rng.integers(0, 2, size=n)
        else:                                               # all equal except a few
            scores = np.zeros(n)
            scores[rng.integers(0, n, size=5)] = rng.integers(-1, 2, size=5)
        for k in (1, 10, 20):
            if k > n:
                continue
            got, _ = emulate(scores, k, rng, merge_at=int(rng.integers(1, 9)))
            want = exact_topk(scores, k)
            assert np.array_equal(got, want), (trial, kind, k, got, want)
            cases += 1
    assert cases > 1000


def test_piece_dump_volume_matches_the_slot_buffer_model():
    """The slot buffers are sized for ~ 2*K*(1 + ln(pieces / 2K)) dumped pieces per row (x 1.6 head-room + 16,
    score_tc.cu::tc_plan): random rows of the evaluation's length stay below that with merges every 8 pending."""
    rng = np.random.default_rng(1)
    n, k = 82_174, 20
    pieces = (n + PIECE - 1) // PIECE
    expect = 2 * k * (1 + np.log(pieces / (2 * k)))
    cap = int(1.6 * expect) + 16
    worst = 0
    for _ in range(5):
        _, dumped = emulate(rng.standard_normal(n), k, rng, merge_at=8)
        worst = max(worst, dumped)
    assert worst < cap, (worst, cap)
