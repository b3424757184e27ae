import torch, torch.nn.functional as F, time
from torch.nn.attention import sdpa_kernel, SDPBackend
torch.manual_seed(0)
B,h,L,D=64,8,1024,64
q=torch.randn(B,h,L,D,device='cuda'); k=torch.randn(B,h,L+1,D,device='cuda'); v=torch.randn(B,h,L+1,D,device='cuda')
def t(fn,n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): y=fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n, y
ms0,y0=t(lambda: F.scaled_dot_product_attention(q,k,v))
print('default', ms0)
for name,b in [('math',SDPBackend.MATH),('efficient',SDPBackend.EFFICIENT_ATTENTION),('flash',SDPBackend.FLASH_ATTENTION),('cudnn',SDPBackend.CUDNN_ATTENTION)]:
    try:
        with sdpa_kernel([b]):
            ms,y=t(lambda: F.scaled_dot_product_attention(q,k,v))
        print(name, ms, 'maxrel', ((y-y0).abs().max()/y0.abs().max()).item())
    except Exception as e:
        print(name,'ERR',str(e)[:100])
# non-contiguous k as in the module (cat of expanded null kv): same
