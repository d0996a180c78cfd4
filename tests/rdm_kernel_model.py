"""TEST INFRASTRUCTURE: host models of the RDM kernels of forte2_b200/csrc/rdm_kernels.cuh.  Each function follows
its kernel statement by statement -- same device-format tables (exported by f2_strings_export kinds 20..28), same
unpacking masks, same offset arithmetic -- with the CTA / thread loops flattened.  The CPU suite checks these models
against the definitional RDMs (oracle/rdm_port.py, pinned to the reference); what is left for the GPU is the CUDA
mechanics (reductions, atomics, launch shapes)."""
import numpy as np

M64 = (1 << 64) - 1


class Tables:
    def __init__(self, strings):
        u = lambda kind, spin=0: strings.export(kind, spin).astype(np.int64).view(np.uint64)  # noqa: E731
        self.norb = strings.norb
        self.nas, self.nbs = strings.nas, strings.nbs
        self.as_ptr = strings.export("dev_as_ptr", 0)
        self.as_ent = u("dev_as_ent")
        bs = strings.export("dev_bs_ent", 0)
        self.bs_width = int(bs[0])
        self.bs_ent = bs[1:].view(np.uint64)
        self.h2_ptr = [strings.export("dev_h2_ptr", s) for s in (0, 1)]
        self.h2_ent = [u("dev_h2_ent", s) for s in (0, 1)]
        self.a_cls, self.b_cls = strings.export("dev_cls", 0), strings.export("dev_cls", 1)
        self.a_loc, self.b_loc = strings.export("dev_loc", 0), strings.export("dev_loc", 1)
        self.nA_cls, self.nB_cls = strings.export("dev_ncls", 0), strings.export("dev_ncls", 1)
        self.nclsA, self.nclsB = len(self.nA_cls), len(self.nB_cls)
        self.blk_off = strings.export("dev_blk_off", 0)


def rdm1a(t, L, R):
    g = np.zeros(t.norb * t.norb)
    for Ia in range(t.nas):
        cI, lI = int(t.a_cls[Ia]), int(t.a_loc[Ia])
        for e in range(int(t.as_ptr[Ia]), int(t.as_ptr[Ia + 1])):
            ent = int(t.as_ent[e])
            Ja = ent & 0xffffffff
            kl = (ent >> 32) & 0xffff
            cJ, lJ = int(t.a_cls[Ja]), int(t.a_loc[Ja])
            acc = 0.0
            for c in range(t.nclsB):
                oI, oJ = int(t.blk_off[cI * t.nclsB + c]), int(t.blk_off[cJ * t.nclsB + c])
                nB = int(t.nB_cls[c])
                if oI < 0 or oJ < 0:
                    continue
                acc += L[oI + lI * nB: oI + lI * nB + nB] @ R[oJ + lJ * nB: oJ + lJ * nB + nB]
            g[kl] += -acc if (ent >> 48) & 1 else acc
    return g.reshape(t.norb, t.norb)


def _beta_entries(t, Jb):
    for w in range(t.bs_width):
        ent = int(t.bs_ent[w * t.nbs + Jb])
        if ent == M64:
            break
        yield ent & 0xfffffff, (ent >> 28) & 0xfff, (ent >> 40) & 0xffff, (ent >> 56) & 1


def rdm1b(t, L, R):
    g = np.zeros(t.norb * t.norb)
    for Ia in range(t.nas):
        cI, lI = int(t.a_cls[Ia]), int(t.a_loc[Ia])
        for Jb in range(t.nbs):
            cJ = int(t.b_cls[Jb])
            oR = int(t.blk_off[cI * t.nclsB + cJ])
            if oR < 0:
                continue
            r = R[oR + lI * int(t.nB_cls[cJ]) + int(t.b_loc[Jb])]
            for ai, ci, ij, neg in _beta_entries(t, Jb):
                oL = int(t.blk_off[cI * t.nclsB + ci])
                if oL < 0:
                    continue
                v = L[oL + lI * int(t.nB_cls[ci]) + ai] * r
                g[ij] += -v if neg else v
    return g.reshape(t.norb, t.norb)


def rdm2ab(t, L, R):
    n = t.norb
    out = np.zeros(n ** 4)
    for Ia in range(t.nas):
        cI, lI = int(t.a_cls[Ia]), int(t.a_loc[Ia])
        for e in range(int(t.as_ptr[Ia]), int(t.as_ptr[Ia + 1])):
            aent = int(t.as_ent[e])
            Ja = aent & 0xffffffff
            ux = (aent >> 32) & 0xffff
            u, x = divmod(ux, n)
            aneg = (aent >> 48) & 1
            cJ, lJ = int(t.a_cls[Ja]), int(t.a_loc[Ja])
            tile = np.zeros(n * n)
            for Jb in range(t.nbs):
                cb = int(t.b_cls[Jb])
                oR = int(t.blk_off[cJ * t.nclsB + cb])
                if oR < 0:
                    continue
                r = R[oR + lJ * int(t.nB_cls[cb]) + int(t.b_loc[Jb])]
                for ai, ci, ij, neg in _beta_entries(t, Jb):
                    oL = int(t.blk_off[cI * t.nclsB + ci])
                    if oL < 0:
                        continue
                    v = L[oL + lI * int(t.nB_cls[ci]) + ai] * r
                    tile[ij] += -v if (neg != aneg) else v
            for i in range(n * n):
                if tile[i] != 0.0:
                    v, y = divmod(i, n)
                    out[((u * n + v) * n + x) * n + y] += tile[i]
    return out.reshape(n, n, n, n)


def rdm2ss(t, L, R, spin):
    n = t.norb
    npairs = n * (n - 1) // 2
    out = np.zeros(npairs * npairs)
    ptr, ents = t.h2_ptr[spin], t.h2_ent[spin]
    for K in range(len(ptr) - 1):
        e0, e1 = int(ptr[K]), int(ptr[K + 1])
        m = e1 - e0
        for ab in range(m * m):
            a, b = divmod(ab, m)
            ea, eb = int(ents[e0 + a]), int(ents[e0 + b])
            Sa, Sb = ea & 0xffffffff, eb & 0xffffffff
            acc = 0.0
            if spin == 0:
                ca, cb = int(t.a_cls[Sa]), int(t.a_cls[Sb])
                la, lb = int(t.a_loc[Sa]), int(t.a_loc[Sb])
                for c in range(t.nclsB):
                    oa, ob = int(t.blk_off[ca * t.nclsB + c]), int(t.blk_off[cb * t.nclsB + c])
                    nB = int(t.nB_cls[c])
                    if oa < 0 or ob < 0:
                        continue
                    acc += L[oa + la * nB: oa + la * nB + nB] @ R[ob + lb * nB: ob + lb * nB + nB]
            else:
                ca, cb = int(t.b_cls[Sa]), int(t.b_cls[Sb])
                la, lb = int(t.b_loc[Sa]), int(t.b_loc[Sb])
                na, nb = int(t.nB_cls[ca]), int(t.nB_cls[cb])
                for c in range(t.nclsA):
                    oa, ob = int(t.blk_off[c * t.nclsB + ca]), int(t.blk_off[c * t.nclsB + cb])
                    nA = int(t.nA_cls[c])
                    if oa < 0 or ob < 0:
                        continue
                    i = np.arange(nA)
                    acc += L[oa + i * na + la] @ R[ob + i * nb + lb]
            if acc != 0.0:
                pa, pb = (ea >> 32) & 0xffff, (eb >> 32) & 0xffff
                neg = ((ea >> 48) & 1) != ((eb >> 48) & 1)
                out[pa * npairs + pb] += -acc if neg else acc
    return out.reshape(npairs, npairs)
