// tt_actor_tc.cu -- actor forward on the 5th-generation tensor cores (tcgen05 + TMEM), bf16 operands,
// fp32 accumulation.  (placeholder until the kernel lands: fails loudly, never falls back)
#include "tt_actor.cuh"
#include "tt_common.cuh"

namespace tt {
int actor_pack_tc(tt_actor *, const float *, const float *, cudaStream_t) { return TT_OK; }
int actor_forward_tc(const tt_actor *, const float *, int64_t, int64_t, float *, cudaStream_t) {
    set_error("tt_actor_forward: TT_PREC_BF16 (tcgen05) path is not built in this revision");
    return TT_ERR_INVALID;
}
}  // namespace tt
