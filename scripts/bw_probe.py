import torch
dev='cuda'
flush = torch.empty(256*1024*1024, dtype=torch.uint8, device=dev)
for mb in (324, 1000):
    x = torch.randn(mb*1024*1024//4, device=dev)
    y = torch.empty_like(x)
    for name, fn in (('sum (read only)', lambda: x.sum()), ('copy (read+write)', lambda: y.copy_(x)), ('fill (write only)', lambda: y.fill_(1.0)), ('max (read only)', lambda: x.max())):
        ts=[]
        for i in range(7):
            flush.fill_(1)
            e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            if i>=2: ts.append(e0.elapsed_time(e1))
        t=sum(ts)/len(ts)
        nbytes = x.numel()*4*(2 if 'copy' in name else 1)
        print('%4d MB %-18s %.1f us  %.2f TB/s' % (mb, name, 1e3*t, nbytes/t/1e9))
