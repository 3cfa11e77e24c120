"""Dev: sustained HBM copy bandwidth (torch b.copy_(a), read+write bytes) over several seconds, with SM clock / power
samples — the ceiling a long HBM-bound scan can be compared with when the board runs into its power cap."""
import subprocess, sys, time
import torch

n = 8 * (1 << 30)                       # 8 GiB per buffer
a = torch.empty(n, dtype=torch.uint8, device="cuda")
b = torch.empty(n, dtype=torch.uint8, device="cuda")
a.fill_(1)
for _ in range(3):
    b.copy_(a)
torch.cuda.synchronize()
smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,clocks_event_reasons.sw_power_cap",
                        "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
secs = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
res = []
t_end = time.time() + secs
while time.time() < t_end:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        b.copy_(a)
    e1.record()
    torch.cuda.synchronize()
    res.append(20 * 2 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9)
smi.terminate()
rows = [l.strip() for l in smi.stdout.read().splitlines() if l.strip()]
print("copy GB/s per 20-copy window:", " ".join(f"{x:.0f}" for x in res))
print("first", f"{res[0]:.0f}", "last", f"{res[-1]:.0f}", "min", f"{min(res):.0f}")
print("smi samples (sm MHz, mem MHz, W, power cap):", rows[:3], "...", rows[-3:])
