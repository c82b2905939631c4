import ctypes, numpy as np, os, torch
torch.zeros(1).cuda()
L = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'libmbar_pingpong.so'))
out = np.zeros(148, np.int64)
for mode, name in ((0, 'try_wait'), (1, 'try_wait+20us hint'), (2, 'test_wait poll')):
    for nextra in (0, 16):
        rc = L.run_pingpong(mode, 20000, nextra, out.ctypes.data_as(ctypes.c_void_p))
        print('%-20s bystanders=%2d rc=%d  round trip (2 hops) cycles: median %d min %d max %d' % (name, nextra, rc, np.median(out), out.min(), out.max()))
