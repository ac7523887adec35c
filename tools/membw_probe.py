"""Probe of achievable HBM bandwidth vs read:write mix with stock torch kernels (context for the
roofline: the decode path writes 73 % of its bytes)."""
import torch, json
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)/n/1e3
N = 1<<30  # elements
x = torch.empty(N, dtype=torch.bfloat16, device='cuda').normal_()
y = torch.empty_like(x); z = torch.empty_like(x)
res = {}
res['write_only_fill_GBps'] = 2*N/t(lambda: y.fill_(1.0))/1e9
res['write_only_zero_GBps'] = 2*N/t(lambda: y.zero_())/1e9
res['copy_1r1w_GBps'] = 4*N/t(lambda: y.copy_(x))/1e9
res['add_2r1w_GBps'] = 6*N/t(lambda: torch.add(x, y, out=z))/1e9
res['read_only_sum_GBps'] = 2*N/t(lambda: x.sum())/1e9
# 1 read : 2.67 write like the decode (1.5 B in, 4 B out): u8 -> 4 x u8 via repeat_interleave-free expand copy
u = torch.empty(1<<29, dtype=torch.uint8, device='cuda')
o = torch.empty((1<<29, 4), dtype=torch.uint8, device='cuda')
res['expand_1r4w_GBps'] = 5*(1<<29)/t(lambda: o.copy_(u[:, None].expand(-1, 4)))/1e9
o32 = torch.empty(1<<29, dtype=torch.int32, device='cuda')
res['cast_u8_to_i32_1r4w_GBps'] = 5*(1<<29)/t(lambda: o32.copy_(u))/1e9
print(json.dumps(res, indent=1))
