#!/usr/bin/env python
"""Can a CTA order cut the DRAM traffic of t3g_kernel?  (round-1 review, item 5a.)  LRU model of the L2: every CTA (input alpha
string Ja of CAS(16,16)) reduces into the 65 sigma rows of its closed single-replacement neighbourhood, 148 CTAs run
together, the cache holds `cap` rows of 103 KB (126 MB of L2 = 1 200 rows if nothing else lived there).  Prints the row
fetches per sigma row for six orders.  Result: 18 - 24 fetches per row whatever the order (lexical, the order used: 19 - 21;
ncu measures 26.3 GB read = 19.8 fetches per row), because the neighbourhoods of the Johnson graph J(16,8) are too
wide for any window of ~1 000 rows; only a random order is worse (57 - 60)."""
import itertools, sys, random
from collections import OrderedDict
n, k = 16, 8
strings = [sum(1 << i for i in c) for c in itertools.combinations(range(n), k)]
strings.sort()
addr = {s: i for i, s in enumerate(strings)}
def neigh(s):
    occ = [i for i in range(n) if s >> i & 1]; emp = [i for i in range(n) if not s >> i & 1]
    out = [addr[s]]
    for i in occ:
        for a in emp:
            out.append(addr[s ^ (1 << i) ^ (1 << a)])
    return out
N = {addr[s]: neigh(s) for s in strings}
def misses(order, cap, conc=148):
    cache = OrderedDict(); miss = 0
    # process in groups of `conc` concurrent CTAs: interleave their touches
    for g0 in range(0, len(order), conc):
        grp = order[g0:g0 + conc]
        for t in range(65):
            for ja in grp:
                r = N[ja][t]
                if r in cache: cache.move_to_end(r)
                else:
                    miss += 1; cache[r] = 1
                    if len(cache) > cap: cache.popitem(last=False)
    return miss
lex = list(range(len(strings)))
# revolving door / gray order: order by reflected gray code of combination
def gray_order():
    def rd(n, k):
        if k == 0: return [0]
        if k == n: return [(1 << n) - 1]
        a = rd(n - 1, k); b = rd(n - 1, k - 1)
        return a + [x | (1 << (n - 1)) for x in reversed(b)]
    return [addr[s] for s in rd(n, k)]
# BFS order from string 0
def bfs():
    from collections import deque
    seen = [False]*len(strings); out=[]; dq=deque([0]); seen[0]=True
    while dq:
        v=dq.popleft(); out.append(v)
        for w in N[v][1:]:
            if not seen[w]: seen[w]=True; dq.append(w)
    return out
# blocked: sort by (low-half pattern) so strings sharing the low 8 bits pattern are consecutive
def by_low():
    return sorted(lex, key=lambda a: (strings[a] & 0xff, strings[a] >> 8))
def by_popcount_low():
    return sorted(lex, key=lambda a: (bin(strings[a] & 0xff).count('1'), strings[a] & 0xff, strings[a] >> 8))
orders = {'lex': lex, 'gray': gray_order(), 'bfs': bfs(), 'by_low': by_low(), 'pop_low': by_popcount_low()}
rnd = lex[:]; random.Random(1).shuffle(rnd); orders['random'] = rnd
for cap in (600, 900, 1200):
    print('cap', cap, {k: round(misses(v, cap) / len(strings), 2) for k, v in orders.items()})
