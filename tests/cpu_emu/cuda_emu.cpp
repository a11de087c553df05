// TEST INFRASTRUCTURE ONLY -- see cuda_emu.h.
#include "cuda_emu.h"

namespace ndp_emu {

thread_local Ctx ctx;
Barrier block_barrier;
Barrier named_barriers[16];
std::vector<Barrier*> warp_barriers;
std::vector<uint64_t> warp_slots;
unsigned char* dyn_smem_ptr = nullptr;
float tmem[128][512];
static size_t dyn_smem_cap = 0;

void prepare(unsigned nthreads, size_t smem) {
    block_barrier.name = "__syncthreads / block end";
    for (int i = 0; i < 16; ++i) named_barriers[i].name = "named barrier";
    block_barrier.reset((int)nthreads);
    unsigned nwarps = (nthreads + 31) / 32;
    while (warp_barriers.size() < nwarps) { warp_barriers.push_back(new Barrier()); warp_barriers.back()->name = "warp barrier"; }
    for (unsigned w = 0; w < nwarps; ++w) {
        unsigned lanes = (w + 1) * 32 <= nthreads ? 32 : nthreads - w * 32;
        warp_barriers[w]->reset((int)lanes);
    }
    warp_slots.assign((size_t)nwarps * 32, 0);
    if (smem > dyn_smem_cap) {
        free(dyn_smem_ptr);
        dyn_smem_cap = (smem + 1023) / 1024 * 1024;
        dyn_smem_ptr = (unsigned char*)aligned_alloc(1024, dyn_smem_cap);
    }
    if (dyn_smem_ptr) memset(dyn_smem_ptr, 0xFF, dyn_smem_cap);   // NaN poison: catch reads of unwritten smem
}

uint64_t warp_exchange(uint64_t v, int src_lane) {
    Barrier* b = warp_barriers[ctx.warp];
    warp_slots[(size_t)ctx.warp * 32 + ctx.lane] = v;
    b->wait();
    uint64_t r = warp_slots[(size_t)ctx.warp * 32 + (src_lane & 31)];
    b->wait();
    return r;
}

unsigned warp_ballot(int pred) {
    Barrier* b = warp_barriers[ctx.warp];
    warp_slots[(size_t)ctx.warp * 32 + ctx.lane] = pred ? 1 : 0;
    b->wait();
    unsigned m = 0;
    for (int l = 0; l < 32; ++l) m |= (unsigned)(warp_slots[(size_t)ctx.warp * 32 + l] & 1) << l;
    b->wait();
    return m;
}

}  // namespace ndp_emu
