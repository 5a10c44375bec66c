"""ALU throughput probe (NVRTC, runs on the GPU box): cycles per warp-instruction per SMSP for the epilogue's
candidate instructions.  One CTA per SM, W warps; each thread runs 16 independent dependency chains of one op."""
import sys, torch
TEMPLATE = r'''
extern "C" __global__ void probe(int iters, long long* out, unsigned* sink) {
  unsigned r[16];
  float f[16];
  for (int i = 0; i < 16; ++i) { r[i] = threadIdx.x * 2654435761u + i; f[i] = __uint_as_float((r[i] & 0x007fffffu) | 0x3f800000u); }
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int j = (i + 1) & 15;
      OP
    }
  }
  long long t1 = clock64();
  unsigned acc = 0;
  for (int i = 0; i < 16; ++i) acc ^= r[i] ^ __float_as_uint(f[i]);
  if (acc == 0x12345678u) sink[0] = acc;
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}
'''
OPS = [
    ("cvt.rn.bf16x2.f32 (F2FP)", 'unsigned d; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(f[i]), "f"(f[j])); f[i] = __uint_as_float(d);', 1),
    ("cvt.rn.relu.bf16x2.f32", 'unsigned d; asm volatile("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(f[i]), "f"(f[j])); f[i] = __uint_as_float(d);', 1),
    ("cvt.rn.f16x2.f32", 'unsigned d; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(f[i]), "f"(f[j])); f[i] = __uint_as_float(d);', 1),
    ("FADD", 'asm volatile("add.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(f[j]));', 1),
    ("FMNMX", 'asm volatile("max.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(f[j]));', 1),
    ("LOP3 (and)", 'asm volatile("and.b32 %0, %0, %1;" : "+r"(r[i]) : "r"(r[j]));', 1),
    ("PRMT", 'asm volatile("prmt.b32 %0, %0, %1, 0x7632;" : "+r"(r[i]) : "r"(r[j]));', 1),
    ("SHL", 'asm volatile("shl.b32 %0, %0, 1;" : "+r"(r[i]));', 1),
    ("IADD", 'asm volatile("add.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(r[j]));', 1),
    ("FFMA", 'asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[i]) : "f"(f[j]));', 1),
]
iters = 1000
for name, op, n in OPS:
    k = torch.cuda._compile_kernel(TEMPLATE.replace("OP", op), "probe", compute_capability="100a")
    line = f"{name:28s}:"
    for warps in (4, 8, 16):
        out = torch.zeros(148, dtype=torch.int64, device="cuda")
        sink = torch.zeros(4, dtype=torch.int32, device="cuda")
        for _ in range(2):
            k(grid=(148, 1, 1), block=(32 * warps, 1, 1), args=[iters, out, sink])
        torch.cuda.synchronize()
        cyc = out.float().mean().item()
        per = cyc / (iters * 16 * n * warps / 4)      # cycles per warp-instruction per SMSP
        line += f"  {warps:2d} warps: {per:5.2f} cyc/warp-inst/SMSP"
    print(line, flush=True)
