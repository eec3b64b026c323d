import sys, os
sys.path[:0]=['/root/repo','/root/repo/multimoda-rs_b200']
os.environ["MMRS_TC_TRACE"]="/root/repo/gpurun_out/tc_trace.txt"
import numpy as np
from multimodars import _native as nat
from scripts.quick_bench import contour
ctx=nat.Context(0)
rng=np.random.default_rng(0); U=4; N=520
t=np.concatenate([contour(rng,N,rng.normal(0,.2)) for _ in range(U)]); r=np.concatenate([contour(rng,N) for _ in range(U)])
off=np.arange(U+1)*N; g=nat.make_grid(0.05,180.0)
ctx.sweep_upload(t,off,r,off,np.full((U,2),4.5),[g],mode=0,prefilter=2)
ctx.sweep_run(); ctx.sweep_download(); print(ctx.timings(), ctx.prefilter_info())
