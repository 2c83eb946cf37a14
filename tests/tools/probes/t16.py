import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
from mla_b200 import _lib
L = _lib.lib(); st = lambda: torch.cuda.current_stream().cuda_stream
N,H,W,Cin,Cout,R,stride,pad = 2,8,8,64,64,1,1,0
x16 = torch.randn(N,H,W,Cin, device="cuda").half(); w16 = torch.randn(Cout,R,R,Cin, device="cuda").half()
y = torch.empty(N,H,W,Cout, device="cuda")
print("fprop16 rc", L.mla_conv2d_fprop16(x16.data_ptr(), w16.data_ptr(), y.data_ptr(), N,H,W,Cin,Cout,R,R,stride,pad,None,st())); 
try:
    torch.cuda.synchronize(); print("fprop16 ok", float((y - (x16.float().view(-1,Cin) @ w16.float().view(Cout,Cin).t()).view_as(y)).abs().max()))
except Exception as e: print("fprop16 FAILED", str(e)[:80]); sys.exit(1)
dy16 = torch.randn(N,H,W,Cout, device="cuda").bfloat16(); wt16 = w16.permute(3,1,2,0).contiguous(); dx = torch.empty(N,H,W,Cin, device="cuda")
print("dgrad16 rc", L.mla_conv2d_dgrad16(dy16.data_ptr(), wt16.data_ptr(), dx.data_ptr(), N,H,W,Cin,Cout,R,R,stride,pad,0,st()))
try:
    torch.cuda.synchronize(); print("dgrad16 ok", float((dx - (dy16.float().view(-1,Cout) @ w16.float().view(Cout,Cin)).view_as(dx)).abs().max()))
except Exception as e: print("dgrad16 FAILED", str(e)[:80])
