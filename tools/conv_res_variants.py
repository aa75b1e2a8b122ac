import torch, torch.nn.functional as F
torch.backends.cudnn.benchmark = True
dev='cuda'; N=40960
x = torch.randn(N,64,11,11,device=dev,dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
z = torch.randn_like(x)
w = (torch.randn(64,64,3,3,device=dev,dtype=torch.bfloat16)*0.05).contiguous(memory_format=torch.channels_last)
b = torch.randn(64,device=dev,dtype=torch.bfloat16)
one,pad=(1,1),(1,1)
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a,c=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    c.record(); torch.cuda.synchronize()
    return a.elapsed_time(c)/reps
for limit in (10, 0):
    torch.backends.cudnn.benchmark_limit = limit
    # new shapes force re-benchmark: perturb by cloning weights
    w2 = w.clone()
    print('limit', limit, 'conv_relu', timeit(lambda: torch.cudnn_convolution_relu(x,w2,b,one,pad,one,1)),
          'conv_add_relu', timeit(lambda: torch.cudnn_convolution_add_relu(x,w2,z,1.0,b,one,pad,one,1)),
          'conv only', timeit(lambda: F.conv2d(x,w2,None,padding=1)),
          'conv+bias', timeit(lambda: F.conv2d(x,w2,b,padding=1)), flush=True)
# alternative: residual folded in as an extra accumulate pass: relu(conv(x)+b) then add? (not equivalent) -- skip
# add+relu elementwise cost
y = torch.empty_like(x)
print('add_relu elementwise', timeit(lambda: torch.relu_(torch.add(x, z, out=y))))
# fp32 residual? alpha=1 with z in same dtype only
