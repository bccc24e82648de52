/* Offline stand-in for cpu_features' cpuinfo_x86.h (see cpu_features_macros.h in this dir).
 * Only the feature bits the reference's dispatch ladders read are modelled
 * (deps/VectorSimilarity/src/VecSim/spaces/{L2_space,IP_space}.cpp). */
#pragma once
namespace cpu_features {
struct X86Features {
    int sse, sse3, sse4_1, avx, avx2, fma3, f16c;
    int avx512f, avx512bw, avx512vl, avx512vnni, avx512vbmi2, avx512dq;
    int avx512_bf16, avx512_fp16;
};
struct X86Info {
    X86Features features;
};
/* Probes the running CPU; bits named in the env var REF_CPU_MASK (comma separated, e.g.
 * "avx512_fp16,avx512f") are forced OFF so a test can pin a lower dispatch tier. */
X86Info GetX86Info();
} // namespace cpu_features
