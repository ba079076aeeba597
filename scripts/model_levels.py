"""Critical-path model of one PCG iteration's triangular solves as a function of the number of nested-dissection leaves T,
from the per-chunk rates measured by bench.py at 256^3 / T = 8 (profiles/r01_bench_lap3d256_T8.json, `tree_levels`):
leaf chain 2.46 chunks/us, separator chains 1.0 (forward) and 1.9 (backward) chunks/us, one launch per level.
Separator at depth d of an n^3 box: ~ (n^3 / 2^d)^(2/3) rows (area of a cut through a sub-box of that volume).
Usage: python scripts/model_levels.py     - prints ms per iteration (leaf level x2, separators forward, backward, total)."""
import math


def model(n, T, leaf_rate=2.46, sep_rate_f=1.0, sep_rate_b=1.9, sms=148, launch_us=8.0):
    N = n ** 3
    k = int(math.log2(T))
    sep_f = sep_b = 0.0
    for d in range(k):
        chunks = (N / 2 ** d) ** (2 / 3) / 32
        waves = math.ceil(2 ** d / sms)
        sep_f += waves * chunks / sep_rate_f / 1e3 + launch_us / 1e3
        sep_b += waves * chunks / sep_rate_b / 1e3 + launch_us / 1e3
    leaf = math.ceil(T / sms) * (N / T / 32) / leaf_rate / 1e3
    return leaf, sep_f, sep_b, 2 * leaf + sep_f + sep_b


if __name__ == "__main__":
    print("n    T     leaf(ms, one direction)  separators fwd  bwd   total ms/iteration")
    for n in (256, 512):
        for T in (8, 64, 256, 1024, 4096):
            print("%-4d %-5d %8.2f %22.2f %6.2f %8.2f" % ((n, T) + model(n, T)))
