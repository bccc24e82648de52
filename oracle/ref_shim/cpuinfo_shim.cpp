// Runtime CPU probe behind the cpu_features stand-in (test infrastructure only).
#include "cpuinfo_x86.h"
#include <cstdlib>
#include <cstring>

namespace cpu_features {
static bool masked(const char *name) {
    const char *m = getenv("REF_CPU_MASK");
    if (!m) return false;
    size_t n = strlen(name);
    for (const char *p = m; *p;) {
        const char *e = strchr(p, ',');
        size_t len = e ? (size_t)(e - p) : strlen(p);
        if (len == n && !strncmp(p, name, n)) return true;
        p += len + (e ? 1 : 0);
    }
    return false;
}
#define PROBE(field, gccname) f.field = !masked(#field) && __builtin_cpu_supports(gccname)
X86Info GetX86Info() {
    __builtin_cpu_init();
    X86Features f{};
    PROBE(sse, "sse");
    PROBE(sse3, "sse3");
    PROBE(sse4_1, "sse4.1");
    PROBE(avx, "avx");
    PROBE(avx2, "avx2");
    PROBE(fma3, "fma");
    PROBE(f16c, "f16c");
    PROBE(avx512f, "avx512f");
    PROBE(avx512bw, "avx512bw");
    PROBE(avx512vl, "avx512vl");
    PROBE(avx512vnni, "avx512vnni");
    PROBE(avx512vbmi2, "avx512vbmi2");
    PROBE(avx512dq, "avx512dq");
    PROBE(avx512_bf16, "avx512bf16");
    PROBE(avx512_fp16, "avx512fp16");
    return X86Info{f};
}
} // namespace cpu_features
