// ORACLE — test infrastructure only (see ora_math.h).  VCM/BDPT programs (placeholder until restated).
#include "ora_render.h"
#include <stdexcept>
namespace ora {
struct LightVertex { int dummy; };
void vcm_lightPass(Renderer&, uint, uint) { throw std::runtime_error("oracle VCM not built yet"); }
void vcm_cameraPass(Renderer&, uint, uint) { throw std::runtime_error("oracle VCM not built yet"); }
void vcm_alloc_vertices(Renderer&, uint) {}
void vcm_free(Renderer&) {}
}
