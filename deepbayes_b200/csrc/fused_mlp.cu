// placeholder: replaced by the tcgen05 fused MLP kernel
#include "engine.h"
namespace dbx {
int fused_mlp_run(dbx_model*, const float*, int64_t, int64_t, const Source&, const Sink&, cudaStream_t) { return 0; }
void fused_mlp_free(dbx_model*) {}
}
