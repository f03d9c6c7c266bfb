// Chains: a recorded sequence of copy / grouped-GEMM launches (several tensordots in a row: the effective-Hamiltonian
// application of DMRG, an environment update) replayed by ONE call.
//
// The reference walks such a sequence through yastn.tensordot four times per application (yastn/tn/mps/_env.py:512-518 Heff2,
// :496-504 update_env_*), re-deriving the block metadata of every contraction in Python each time although the block structure
// of the operands has not changed since the previous Lanczos iteration.  A chain keeps the plans of the first run and the data
// flow between them (slots = the caller's operands, one scratch arena for the intermediates, the result); replaying it costs
// one library call and the kernel launches — no metadata pass, no per-step allocation.  Launches go to the caller's stream in
// recorded order and are CUDA-graph capturable like the individual runs.
#include "yb_common.h"

using namespace yb;

struct yb_chain {
    struct Step {
        int kind, flags;
        const void* plan;
        int slot[3];
        int64_t off[3];
        int64_t dst_elems;
    };
    std::vector<Step> steps;
    int nslots = 0;
};

extern "C" int yb_chain_create(const int64_t* steps, int64_t nsteps, int64_t nslots, yb_chain** out) {
    if (!out) return fail(kErrArg, "yb_chain_create: out is null");
    *out = nullptr;
    if (nsteps < 0 || nslots <= 0 || (nsteps > 0 && !steps)) return fail(kErrArg, "yb_chain_create: bad arguments");
    auto* ch = new yb_chain();
    ch->nslots = (int)nslots;
    ch->steps.resize((size_t)nsteps);
    for (int64_t i = 0; i < nsteps; ++i) {
        const int64_t* r = steps + i * 10;
        yb_chain::Step& s = ch->steps[(size_t)i];
        s.kind = (int)r[0];
        s.flags = (int)r[1];
        s.plan = (const void*)(uintptr_t)r[2];
        for (int k = 0; k < 3; ++k) {
            s.slot[k] = (int)r[3 + 2 * k];
            s.off[k] = r[4 + 2 * k];
        }
        s.dst_elems = r[9];
        const bool gemm = s.kind == YB_CHAIN_GEMM;
        if ((s.kind != YB_CHAIN_COPY && !gemm) || !s.plan || s.slot[0] < 0 || s.slot[0] >= nslots || s.slot[2] < 0 || s.slot[2] >= nslots ||
            (gemm && (s.slot[1] < 0 || s.slot[1] >= nslots))) {
            delete ch;
            return fail(kErrArg, "yb_chain_create: step %lld is malformed", (long long)i);
        }
    }
    *out = ch;
    return kOk;
}

extern "C" int yb_chain_run(const yb_chain* chain, void* const* slots, int64_t nslots, void* stream) {
    if (!chain || !slots) return fail(kErrArg, "yb_chain_run: null argument");
    if (nslots != chain->nslots) return fail(kErrArg, "yb_chain_run: %lld slots given, the chain has %d", (long long)nslots, chain->nslots);
    for (const auto& s : chain->steps) {
        char* a = (char*)slots[s.slot[0]] + s.off[0];
        char* c = (char*)slots[s.slot[2]] + s.off[2];
        int rc;
        if (s.kind == YB_CHAIN_COPY) {
            rc = yb_copy_run((const yb_copy_plan*)s.plan, a, c, s.dst_elems, s.flags, stream);
        } else {
            char* b = (char*)slots[s.slot[1]] + s.off[1];
            rc = yb_gemm_run((const yb_gemm_plan*)s.plan, a, b, c, s.flags, stream);
        }
        if (rc != kOk) return rc;
    }
    return kOk;
}

extern "C" int64_t yb_chain_steps(const yb_chain* chain) { return chain ? (int64_t)chain->steps.size() : 0; }

extern "C" void yb_chain_destroy(yb_chain* chain) { delete chain; }
