"""Debug: pipeline timeline of CTA 0 of the tensor-core NN GEMM (clock64 at barrier events)."""
import ctypes, sys, torch
sys.path.insert(0, '.')
from fieldconv_b200 import _lib, ops
lib = _lib.load()
lib.fcb_debug_trace.argtypes = [ctypes.c_void_p]
dev = 'cuda:0'
M, N, K = int(sys.argv[1]) if len(sys.argv) > 1 else 80656, 96, 2880
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 1
trans = len(sys.argv) > 3 and sys.argv[3] == 'tn'
if trans:
    a = torch.randn(M, K, device=dev); b = torch.randn(M, N, device=dev)      # P[K x N] = a^T b, reduction over M rows
else:
    a = torch.randn(M, K, device=dev); b = torch.randn(K, N, device=dev)
for _ in range(3): ops.gemm(a, b, trans, mode)
buf = torch.zeros(8 * 64, dtype=torch.int64, device=dev)
lib.fcb_debug_trace(buf.data_ptr())
ops.gemm(a, b, trans, mode); torch.cuda.synchronize()
lib.fcb_debug_trace(None)
t = buf.cpu().view(8, 64)
t0 = int(t[0, 0])
names = ['P:before wait empty', 'P:after wait empty', 'P:arrived full_a', 'M:before wait full_a', 'M:after full_a', 'M:after full_b', 'M:issued+commit', 'B:after wait empty']
print('chunk ' + ' '.join('%9s' % n.split(':')[0] + str(i) for i, n in enumerate(names)))
for kc in range(40):
    print('%5d ' % kc + ' '.join('%10d' % (int(t[i, kc]) - t0) for i in range(8)))
for i, n in enumerate(names): print(i, n)
