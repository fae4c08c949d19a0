"""How much of the batch tail could work-queue ORDER recover?  (CPU study with the oracle; no GPU needed.)

The persistent kernel hands soundings to warps in index order.  Chain lengths differ (10 000 iterations for a chain
that never burns in, 10 001 + burn-in iteration otherwise, plus the iterations before a reset()), so the makespan of
a batch is that of list scheduling on `slots` = SMs x resident chains.  This script runs the oracle on the first N
synthetic soundings of BASELINE configs[1], records each chain's total iterations and what is known about it at
start-up, and simulates the makespan (in iteration-times, all slots equally fast, no speculation) for several orders:
index order (what the kernel does), longest-first with perfect knowledge (the bound for any ordering), and
longest-first by a start-up predictor.
"""
import heapq
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle_py as O  # noqa: E402
from geobipy_b200.synthetic import synthetic_sounding  # noqa: E402


def makespan(lengths, order, slots):
    free = [0.0] * slots
    heapq.heapify(free)
    end = 0.0
    for i in order:
        t = heapq.heappop(free) + lengths[i]
        end = max(end, t)
        heapq.heappush(free, t)
    return end


def main(n=1024, out=None):
    O.build()
    s = O.make_system()
    o = O.resolve_options(n_markov_chains=10000)

    def run(i):
        edges, sigma, z, noise = synthetic_sounding(i)
        clean = O.fdem_forward(s, z, sigma, np.diff(edges))
        data = clean + noise * np.sqrt((0.05 * clean) ** 2 + 25.0)
        r = O.run_chain(s, o, data, z, 0, i)
        sc = r["scalars"]
        return (sc[O.S_TOTAL_ITER], sc[O.S_BURNED_IN], sc[O.S_BURNED_IN_ITER], sc[O.S_N_RESETS], r["misfit_trace"][0],
                float(np.abs(data).max()), z, sigma.size)
    t0 = time.time()
    with ThreadPoolExecutor(os.cpu_count() or 1) as ex:
        rows = np.array(list(ex.map(run, range(n))))
    total, burned, burn_it, resets, misfit0, amp, z, ntrue = rows.T
    slots = int(round(n * 2368 / 4096))       # the bench's ratio of soundings to resident chains (148 SMs x 16)
    idx = np.arange(n)
    res = dict(n=n, slots=slots, seconds=time.time() - t0, total_iterations=float(total.sum()),
               ideal=float(total.sum() / slots), longest=float(total.max()),
               burned_in_fraction=float(burned.mean()), with_resets_fraction=float((resets > 0).mean()),
               length_percentiles={p: float(np.percentile(total, p)) for p in (5, 25, 50, 75, 95, 100)})
    res["makespan_index_order"] = makespan(total, idx, slots)
    res["makespan_longest_first_oracle"] = makespan(total, np.argsort(-total), slots)
    res["makespan_shortest_first_oracle"] = makespan(total, np.argsort(total), slots)
    rng = np.random.default_rng(0)
    res["makespan_random_order_mean"] = float(np.mean([makespan(total, rng.permutation(n), slots) for _ in range(20)]))
    preds = {"initial_misfit": misfit0, "log_initial_misfit": np.log(misfit0), "max_amplitude": amp, "height": z}
    res["predictors"] = {}
    for name, p in preds.items():
        rk = np.corrcoef(np.argsort(np.argsort(p)), np.argsort(np.argsort(total)))[0, 1]
        res["predictors"][name] = dict(rank_correlation=float(rk),
                                       makespan_high_first=makespan(total, np.argsort(-p), slots),
                                       makespan_low_first=makespan(total, np.argsort(p), slots))
    print(json.dumps(res, indent=1))
    if out:
        json.dump(res, open(out, "w"), indent=1)


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 1024, sys.argv[2] if len(sys.argv) > 2 else None)
