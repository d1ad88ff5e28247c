// bootstrap.cu -- placeholder
#include "engine.cuh"
namespace ckks {
struct BootPlan { int out_level = 0; };
void Engine::bootstrap_setup() { throw std::runtime_error("bootstrap: not built yet"); }
Ct* Engine::bootstrap(Ct*) { throw std::runtime_error("bootstrap: not built yet"); }
int Engine::boot_out_level() const { return boot ? boot->out_level : -1; }
Ct* Engine::mod_raise(Ct*) { throw std::runtime_error("bootstrap: not built yet"); }
}  // namespace ckks
