// vcm.cu — vertex-connection BDPT passes (placeholder until the kernels land)
#include "renderer.h"
namespace opr {
struct VcmState { int dummy; };
void vcm_free(Renderer& r) { delete r.vcm; r.vcm = nullptr; }
void Renderer::renderVCM(unsigned long long, unsigned, float, const opr_request_details&)
{
    throw std::runtime_error("VCM_BIDIRECTIONAL_PATH_TRACING kernels are not built yet");
}
}
