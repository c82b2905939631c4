import numpy as np, sys
t = np.fromfile(sys.argv[1], dtype=np.int64).reshape(-1, 8, 64)
t0 = t[t > 0].min()
def rel(a): return np.where(a > 0, a - t0, -1)
mma = rel(t[0, :4])            # [g][sub]: time the MMA thread saw stage g full
f0 = rel(t[1]); f3 = rel(t[3])  # former warp 0 (g=0: stages 0,2) and warp 12 (g=3: stages 1,3): events [4h + {0 before wait, 1 after wait, 2 after arrive}]
print('sub | MMA sees full[0..3]            | warp0: h0 wait_start, wait_end, arrive | h1 ... | warp12: h0.. h1..')
for c in range(8, 24):
    print(c, mma[:, c], '|', f0[0:3, c], f0[4:7, c], '|', f3[0:3, c], f3[4:7, c])
print('MMA per-sub-chunk period (cycles):', np.diff(mma[0, 8:40]).mean(), ' stage-to-stage gaps within a sub-chunk:', (mma[1:, 8:40] - mma[:-1, 8:40]).mean(axis=1))
print('former warp0: wait durations h0 / h1:', (f0[1] - f0[0])[8:40].mean(), (f0[5] - f0[4])[8:40].mean(), ' work h0/h1:', (f0[2] - f0[1])[8:40].mean(), (f0[6] - f0[5])[8:40].mean())

mw = rel(t[0, 4:8])
b = rel(t[2])
print('MMA thread: wait duration per stage (cycles):', (mma - mw)[:, 8:40].mean(axis=1), ' busy between stages (wait_end -> next wait_start):', (mw[1:, 8:40] - mma[:-1, 8:40]).mean(axis=1))
print('B warp (stage 0): wait %.0f  cells %.0f  fence+arrive %.0f ; arrives %.0f cycles before MMA sees full[0]' % ((b[1]-b[0])[8:40].mean(), (b[2]-b[1])[8:40].mean(), (b[3]-b[2])[8:40].mean(), (mma[0]-b[3])[8:40].mean()))
print('former warp0 stage0 arrives %.0f cycles before MMA sees full[0]' % (mma[0]-f0[2])[8:40].mean())
print('former warp0 h0: form+issue STTM %.0f, wait::st+fence+arrive %.0f | h1: %.0f, %.0f' % ((f0[3]-f0[1])[8:40].mean(), (f0[2]-f0[3])[8:40].mean(), (f0[7]-f0[5])[8:40].mean(), (f0[6]-f0[7])[8:40].mean()))

if t.shape[0] > 4 and (t[4] > 0).any():
    fine = rel(t[4])       # warp 0, h=0: [1] I_i loads landed, [2] tile0 math done, [3] tile0 STTM issued, [4] tile1 math done, [5] tile1 STTM issued, [6] wait::st done
    w_end = f0[1]
    sl = slice(8, 40)
    print('former warp0 h0 fine: wait_end->I_i landed %.0f | ->tile0 loads+math %.0f | STTM0 issue %.0f | tile1 loads+math %.0f | STTM1 issue %.0f | wait::st %.0f | fence+arrive %.0f' % (
        (fine[1] - w_end)[sl].mean(), (fine[2] - fine[1])[sl].mean(), (fine[3] - fine[2])[sl].mean(), (fine[4] - fine[3])[sl].mean(),
        (fine[5] - fine[4])[sl].mean(), (fine[6] - fine[5])[sl].mean(), (f0[2] - fine[6])[sl].mean()))
