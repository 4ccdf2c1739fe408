"""CPU model behind the LB2_ATTN_ROTATE experiment (attention_tc.cu): exponentials per SM sub-partition of the tcgen05
attention kernel for a passage-length mix.  Softmax warp w of a CTA owns MMA rows 32w..32w+31 and runs on sub-partition w;
an item (passage, 128-row query block, head) costs every ACTIVE warp Lp = ceil16(L) key columns.  Prints the per-quarter
load (relative to the mean) and the critical quarter's load per token for: no rotation, uniform rotation, the tables in the
kernel.  The model predicted 0.79 -> 0.62 ms on N(128, 48); the measurement (profiles/r02e_attention_rotation_ab.log) shows
no change, i.e. the kernel is not bound by per-sub-partition MUFU throughput but by the per-item chain of a CTA."""
import numpy as np

ROT1 = [3, 3, 3, 3, 3, 3, 0, 1]
ROT2 = [2, 2, 2, 2, 2, 0, 0, 0]


def loads(L, t1, t2, t3):
    n = len(L)
    Lp = (L + 15) // 16 * 16
    h = ((np.arange(n).astype(np.uint64) * 2654435761) % (1 << 32)) >> 13
    ld = np.zeros(4)
    wt0 = (np.minimum(L, 128) + 31) // 32
    for w in range(1, 5):
        ld[:w] += Lp[wt0 == w].sum()
    tail = L > 128
    wt = (L - 128 + 31) // 32
    for w, T in ((1, t1), (2, t2), (3, t3), (4, [0])):
        m = tail & (wt == w)
        rot = np.array(T)[(h[m] % len(T)).astype(int)]
        for r in np.unique(rot):
            ld[r:r + w] += Lp[m][rot == r].sum()
    return ld


if __name__ == "__main__":
    rng = np.random.default_rng(1)
    for mu, sd in ((128, 48), (96, 40), (160, 40), (200, 30)):
        L = np.clip(np.round(rng.normal(mu, sd, 300000)), 16, 256).astype(int)
        for name, T in (("none", ([0], [0], [0])), ("uniform", ([0, 1, 2, 3], [0, 2], [0, 1])), ("kernel tables", (ROT1, ROT2, [0]))):
            ld = loads(L, *T)
            print(f"N({mu},{sd}) {name:14s} quarter load / mean {np.round(ld / ld.mean(), 3)}  critical quarter per token {ld.max() / L.sum():.3f}")
