import torch, time
import torch.nn.functional as F
torch.backends.cudnn.benchmark = True
N=65536
x8 = torch.randint(0,255,(N,4,84,84),dtype=torch.uint8,device='cuda')
def timeit(f, n=5):
    f(); torch.cuda.synchronize()
    s=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): f()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e)/n
w = torch.randn(16,4,8,8,device='cuda')
# 1: NCHW bf16
wb = w.bfloat16()
print("to bf16 + div NCHW: %.2f ms" % timeit(lambda: x8.to(torch.bfloat16)/255.0))
xb = (x8.to(torch.bfloat16)/255.0)
print("conv1 NCHW bf16: %.2f ms" % timeit(lambda: F.conv2d(xb, wb, stride=4)))
xcl = xb.contiguous(memory_format=torch.channels_last); wcl = wb.contiguous(memory_format=torch.channels_last)
print("conv1 NHWC bf16 C=4: %.2f ms" % timeit(lambda: F.conv2d(xcl, wcl, stride=4)))
# 2: pad to 8 channels NHWC
xp = torch.zeros(N,8,84,84,dtype=torch.bfloat16,device='cuda').contiguous(memory_format=torch.channels_last); xp[:, :4] = xb
wp = torch.zeros(16,8,8,8,dtype=torch.bfloat16,device='cuda'); wp[:, :4] = wb; wp = wp.contiguous(memory_format=torch.channels_last)
print("conv1 NHWC bf16 C=8: %.2f ms" % timeit(lambda: F.conv2d(xp, wp, stride=4)))
# 3: fp16
xh = xb.half(); wh = w.half()
print("conv1 NCHW fp16: %.2f ms" % timeit(lambda: F.conv2d(xh, wh, stride=4)))
# 4: space-to-depth: stride-4 8x8 conv == 2x2 conv over 4x4 blocks with 64 channels
xs = xb.view(N,4,21,4,21,4).permute(0,1,3,5,2,4).reshape(N,64,21,21).contiguous(memory_format=torch.channels_last)
ws = wb.view(16,4,2,4,2,4).permute(0,1,3,5,2,4).reshape(16,64,2,2).contiguous(memory_format=torch.channels_last)
print("conv1 as 2x2 over 64ch NHWC: %.2f ms" % timeit(lambda: F.conv2d(xs, ws)))
ref = F.conv2d(xb.float(), w, stride=4); got = F.conv2d(xs, ws).float()
print("s2d max err", (ref-got).abs().max().item())
h1 = F.relu(F.conv2d(xs, ws))
w2 = torch.randn(32,16,4,4,device='cuda').bfloat16().contiguous(memory_format=torch.channels_last)
print("conv2 NHWC: %.2f ms" % timeit(lambda: F.conv2d(h1, w2, stride=2)))
