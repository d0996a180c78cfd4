// TEST INFRASTRUCTURE: the RDM kernels of forte2_b200/csrc/rdm_kernels.cuh executed by the SIMT emulator (simt.h).
// rdm_kernels_host.cuh is generated from the real header by build.sh.
#include "simt.h"
#include "rdm_kernels_host.cuh"

extern "C" int emu_rdm(int kind, int block, const double* L, const double* R, double* out, const uint32_t* a_cls,
                       const uint32_t* a_loc, const uint32_t* b_cls, const uint32_t* b_loc, const uint32_t* nA_cls,
                       const uint32_t* nB_cls, const int64_t* blk_off, const uint32_t* as_ptr, const uint64_t* as_ent,
                       const uint64_t* bs_ent, const uint32_t* h2_ptr, const uint64_t* h2_ent, int nk, int norb,
                       int nclsA, int nclsB, int bs_width, int64_t nas, int64_t nbs) {
    f2::RdmParams p;
    p.L = L;
    p.R = R;
    p.out = out;
    p.a_cls = a_cls;
    p.a_loc = a_loc;
    p.b_cls = b_cls;
    p.b_loc = b_loc;
    p.nA_cls = nA_cls;
    p.nB_cls = nB_cls;
    p.blk_off = blk_off;
    p.as_ptr = as_ptr;
    p.as_ent = as_ent;
    p.bs_ent = bs_ent;
    p.h2_ptr = h2_ptr;
    p.h2_ent = h2_ent;
    p.norb = norb;
    p.nclsA = nclsA;
    p.nclsB = nclsB;
    p.bs_width = bs_width;
    p.nas = nas;
    p.nbs = nbs;
    const int npairs = norb * (norb - 1) / 2;
    switch (kind) {
    case 0: simt::launch(static_cast<unsigned>(nas), block, [&] { f2::rdm1a_kernel(p); }); break;
    case 1: simt::launch(static_cast<unsigned>(nas), block, [&] { f2::rdm1b_kernel(p); }); break;
    case 5: simt::launch(static_cast<unsigned>(nas), block, [&] { f2::rdm2ab_kernel(p); }); break;
    case 3: simt::launch(static_cast<unsigned>(nk), block, [&] { f2::rdm2ss_kernel<0>(p, npairs); }); break;
    case 4: simt::launch(static_cast<unsigned>(nk), block, [&] { f2::rdm2ss_kernel<1>(p, npairs); }); break;
    default: return 1;
    }
    return 0;
}
