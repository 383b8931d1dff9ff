// ORACLE/_ref — test infrastructure only.  Registration interface between the generated translation-unit wrappers
// (one per reference .cu, see gen_tus.py) and the runtime in shim_rt.cpp.
#pragma once
#include "shim/shim_optix.h"
#include <functional>

namespace shim {

enum VarKind { VK_CONTEXT, VK_LAUNCH_INDEX, VK_LAUNCH_DIM, VK_RAY, VK_THIT, VK_PAYLOAD, VK_ATTRIBUTE };
struct VarPlan { Var var; int id; VarKind kind; };
struct BufReg { const char* name; size_t elemSize; void (*bind)(const BufferDesc&); int id; };
struct SamplerReg { const char* name; void (*bind)(const TextureData*); };
struct ProgReg { const char* name; void (*fn)(); void (*fnInt)(int); };
struct ThreadState { const void* ctx = nullptr; unsigned long long epoch = ~0ull; const void* lastInstance = nullptr; };

struct TURuntime {
    TU* tu;
    std::vector<VarPlan> plan;
    std::vector<BufReg> buffers;
    std::vector<SamplerReg> samplers;
    std::vector<ProgReg> programs;
    ThreadState& (*threadStateFn)() = nullptr;
    bool finalized = false;
    explicit TURuntime(TU* t);
    ThreadState& threadState() { return threadStateFn(); }
    size_t saveBytes() const { size_t n = 0; for (const VarPlan& p : plan) n += p.var.size; return n + 16; }
    void (*programOrNull(const char* n))() { for (auto& p : programs) if (!strcmp(p.name, n) && p.fn) return p.fn; return nullptr; }
    void (*program(const char* n))() { auto f = programOrNull(n); if (!f) { fprintf(stderr, "refshim: %s has no program %s\n", tu->name, n); abort(); } return f; }
    void (*intProgram(const char* n))(int) { for (auto& p : programs) if (!strcmp(p.name, n) && p.fnInt) return p.fnInt; fprintf(stderr, "refshim: %s has no program %s\n", tu->name, n); abort(); }
    size_t bufferElemSize(const char* n) const { for (auto& b : buffers) if (!strcmp(b.name, n)) return b.elemSize; return 0; }
};

} // namespace shim
