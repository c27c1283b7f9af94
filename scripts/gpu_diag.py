"""One-shot GPU diagnostics for kernel bring-up: each case runs in its own subprocess under a timeout so that a trap
or a hang in one kernel neither poisons the CUDA context of the others nor stalls the box."""
import subprocess
import sys

CASES = {
    "gemm_small_f32": "from test_kernels_gpu import run_gemm; import torch\n"
    "g,w=run_gemm(128,256,64,4); e=(g-w).abs(); print('max',e.max().item(),'ref',w.abs().max().item());\n"
    "print('row-block errs',[round(e[i*32:(i+1)*32].max().item(),4) for i in range(4)]);\n"
    "print('col-block errs',[round(e[:,i*32:(i+1)*32].max().item(),4) for i in range(8)]);\n"
    "print('got[0,:8]',g[0,:8].tolist()); print('want[0,:8]',w[0,:8].tolist())",
    "gemm_k256_f32": "from test_kernels_gpu import run_gemm\n"
    "g,w=run_gemm(128,256,256,4); e=(g-w).abs(); print('max',e.max().item(),'ref',w.abs().max().item())",
    "gemm_multi_tile": "from test_kernels_gpu import run_gemm\n"
    "g,w=run_gemm(1000,3072,1024,4); e=(g-w).abs(); print('max',e.max().item(),'ref',w.abs().max().item())",
    "gemm_swiglu": "from test_kernels_gpu import run_gemm\n"
    "g,w=run_gemm(256,512,1024,2); e=(g-w).abs(); print('max',e.max().item(),'ref',w.abs().max().item())",
    "attention": "import torch; from test_kernels_gpu import *\n"
    "test_attention_matches_reference(3,14,8,128); print('ok')",
}

if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    for n in names:
        code = "import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')\n" + CASES[n]
        try:
            r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=180)
            print(f"=== {n}: rc={r.returncode}\n{r.stdout[-1500:]}\n{r.stderr[-1500:]}")
        except subprocess.TimeoutExpired:
            print(f"=== {n}: TIMEOUT")
        sys.stdout.flush()
