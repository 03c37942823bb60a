import sys, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fortattack_b200 as fab
L = fab._capi.probe_lib()      # libfortattack_probe.so (test infrastructure)
buf = torch.randn(8 * 16384 // 2, device="cuda").half()
out = torch.zeros(3, dtype=torch.int64, device="cuda"); err = torch.zeros(1, dtype=torch.int32, device="cuda")
for grid in (1, 148):
    for N in (64, 128, 256):
        for depth in (1, 2, 4, 8):
            reps = 64
            rc = L.mp_probe_timing(buf.data_ptr(), N, reps, 16384, depth, out.data_ptr(), err.data_ptr(), grid, None)
            torch.cuda.synchronize()
            o = out.tolist()
            print("grid %3d N=%3d depth=%d: mma %.1f cyc per 128xNx16 (%.0f per 8-kstep chunk) | bulk 16KB sequential %.0f cyc each | depth-%d pipelined %.0f cyc each (%.1f B/cyc)  err %d rc %d"
                  % (grid, N, depth, o[0] / (reps * 8), o[0] / reps, o[1] / reps, depth, o[2] / reps, 16384 * reps / o[2], int(err.item()), rc), flush=True)
