#!/bin/bash
# 1-GPU checks of this session's new kernels + a few timings
set -x
export TD_NO_AUTOBUILD=1 PYTHONPATH=$PWD:$PYTHONPATH
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_gemm_gpu.py -q > gpurun_out/test_ops.log 2>&1; echo "ops rc=$?"; tail -15 gpurun_out/test_ops.log
timeout 120 python - <<'PY' > gpurun_out/quick_timing.log 2>&1
import torch, json, time
import triton_dist.utils as U
U.initialize_distributed(seed=0)
from triton_dist.ops.flash_decode import gqa_fwd_batch_decode
from triton_dist.ops.moe import transposed_moe_grouped_gemm
from triton_dist.ops.gemm import gemm
bf=torch.bfloat16
def timed(fn,n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
res={}
for B,L in ((8,8192),(1,65536),(32,2048),(1,1048576//8)):
    kc=torch.randn(B,L,8,128,device="cuda",dtype=bf); vc=torch.randn_like(kc); q=torch.randn(B,32,128,device="cuda",dtype=bf)
    lens=torch.full((B,),L,device="cuda",dtype=torch.int32)
    ms=timed(lambda: gqa_fwd_batch_decode(q,kc,vc,lens))
    res[f"flash_decode B{B} L{L}"]={"us":round(ms*1e3,1),"kv_TBps":round(2*kc.numel()*2/ms/1e9,3)}
    del kc,vc
# wgrad: Mixtral-like: 16384 rows, 8 experts, N=4096 K=1792
dy=(torch.randn(16384,4096,device="cuda")*0.1).to(bf); x=(torch.randn(16384,1792,device="cuda")*0.1).to(bf)
sp=torch.tensor([2048]*8,device="cuda",dtype=torch.int32)
ms=timed(lambda: transposed_moe_grouped_gemm(dy,x,sp))
res["wgrad 16384x(4096x1792) 8 experts"]={"ms":round(ms,4),"tflops":round(2*16384*4096*1792/ms/1e9,1)}
a=torch.randn(4,4096,device="cuda",dtype=bf); b=torch.randn(12288,4096,device="cuda",dtype=bf)
ms=timed(lambda: gemm(a,b)); res["gemv 4x12288x4096"]={"us":round(ms*1e3,1),"TBps":round(b.numel()*2/ms/1e9,3)}
print(json.dumps(res,indent=1))
PY
cat gpurun_out/quick_timing.log | tail -30
