#!/usr/bin/env python
"""Brute-force the shared-memory pitches (PK = k1 pitch, PV = vector pitch, complex units) of the strip kernels'
exchange buffer X[v][k1][n2] so that the 64-bit accesses of step 1 (stores) and step 2 (loads, in-place
stores) are bank-conflict free for both lane orders (row pass: n2 / k1 fastest; column pass: v fastest)."""
NV = 4
def wavefronts(addrs):
    tot = 0
    for h in range(2):
        lanes = [a for a in addrs[16 * h:16 * h + 16] if a is not None]
        if not lanes: continue
        banks = {}
        for a in set(lanes): banks[a % 16] = banks.get(a % 16, 0) + 1
        tot += max(banks.values())
    return tot
def cost(N1, N2, PK, PV, rowpass):
    c = 0
    t1, t2 = NV * N2, NV * N1
    for r in range((t1 + 31) // 32):
        for k1 in range(N1):
            addrs = []
            for l in range(32):
                task = r * 32 + l
                if task >= t1: addrs.append(None); continue
                if rowpass: v, n2 = divmod(task, N2)
                else: n2, v = divmod(task, NV)
                addrs.append(v * PV + k1 * PK + n2)
            c += wavefronts(addrs)
    for r in range((t2 + 31) // 32):
        for n2 in range(N2):
            addrs = []
            for l in range(32):
                task = r * 32 + l
                if task >= t2: addrs.append(None); continue
                if rowpass: v, k1 = divmod(task, N1)
                else: k1, v = divmod(task, NV)
                addrs.append(v * PV + k1 * PK + n2)
            c += 3 * wavefronts(addrs)      # two loads and one store per element in step 2
    return c
if __name__ == "__main__":
    for (N1, N2) in [(8, 25), (16, 16)]:
        for rowpass in (True, False):
            best = []
            for PK in range(N2, N2 + 8):
                for PV in range(N1 * PK, N1 * PK + 17):
                    best.append((cost(N1, N2, PK, PV, rowpass), PV, PK))
            best.sort()
            print(N1, N2, "row" if rowpass else "col", "(cost, PV, PK):", best[:4])
