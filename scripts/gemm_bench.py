import torch, time, sys
sys.path.insert(0,'/root/repo')
from clip_based_cross_modal_hash_b200 import _lib
lib=_lib.lib()
def run(M,N,K,epi=0,iters=20):
    a=torch.randn(M,K,device='cuda').to(torch.bfloat16); w=torch.randn(N,K,device='cuda').to(torch.bfloat16)
    bias=torch.randn(N,device='cuda'); out=torch.empty(M,N,dtype=torch.bfloat16,device='cuda')
    st=torch.cuda.current_stream().cuda_stream
    for _ in range(3): lib.cmh_gemm_bf16(a.data_ptr(),M,K,K,w.data_ptr(),N,K,bias.data_ptr(),epi,out.data_ptr(),N,None,0,st)
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): lib.cmh_gemm_bf16(a.data_ptr(),M,K,K,w.data_ptr(),N,K,bias.data_ptr(),epi,out.data_ptr(),N,None,0,st)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/iters
    e0.record()
    for _ in range(iters): torch.matmul(a,w.t())
    e1.record(); torch.cuda.synchronize()
    ms2=e0.elapsed_time(e1)/iters
    print('M=%d N=%d K=%d: ours %.1f us %.0f TF/s | cublas %.1f us %.0f TF/s'%(M,N,K,ms*1e3,2*M*N*K/ms/1e9,ms2*1e3,2*M*N*K/ms2/1e9))
for s in [(12800,2304,768),(12800,768,768),(12800,3072,768),(12800,768,3072),(8192,1536,512),(8192,512,512),(8192,2048,512),(8192,512,2048)]: run(*s)
