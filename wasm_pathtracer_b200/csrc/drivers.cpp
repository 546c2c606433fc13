// Photon warm-up, adaptive sampling and the reference-style compute() tick driver.
#include "context.h"

namespace wpt {

uint64_t Context::render_adaptive(uint64_t) { throw std::runtime_error("render_adaptive: not implemented yet"); }
void Context::build_photons() { throw std::runtime_error("build_photons: not implemented yet"); }
void Context::compute(uint64_t) { throw std::runtime_error("compute: not implemented yet"); }
void Context::photon_sample_batch(const float*, const uint32_t*, uint64_t, uint32_t*, float*) { throw std::runtime_error("photon_sample: not implemented yet"); }
void Context::error_map(float*, float*) { throw std::runtime_error("error_map: not implemented yet"); }
void Context::round_spp(uint32_t*) { throw std::runtime_error("round_spp: not implemented yet"); }

}  // namespace wpt
