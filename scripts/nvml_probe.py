import time, torch, pynvml as nv, threading
nv.nvmlInit(); h = nv.nvmlDeviceGetHandleByIndex(0)
x = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
def load():
    for _ in range(300): y = x @ x
    torch.cuda.synchronize()
t = threading.Thread(target=load); t.start()
time.sleep(0.05)
for name, fn in [("clock", lambda: nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)),
                 ("reasons", lambda: nv.nvmlDeviceGetCurrentClocksEventReasons(h)),
                 ("clock", lambda: nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)),
                 ("reasons", lambda: nv.nvmlDeviceGetCurrentClocksEventReasons(h)),
                 ("power", lambda: nv.nvmlDeviceGetPowerUsage(h))]:
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); v = fn(); ts.append(1e3 * (time.perf_counter() - t0))
    print(name, v, ["%.2f" % a for a in ts])
t.join()
# launch latency with / without a polling thread
def launches(n=2000):
    a = torch.zeros(1, device="cuda")
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): a.add_(1)
    torch.cuda.synchronize(); return 1e6 * (time.perf_counter() - t0) / n
print("launch us, no poll", launches())
stop = False
def poll(period, what):
    while not stop:
        if what & 1: nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
        if what & 2: nv.nvmlDeviceGetCurrentClocksEventReasons(h)
        time.sleep(period)
for period, what in [(0.01, 1), (0.01, 2), (0.01, 3), (0.05, 3)]:
    stop = False
    th = threading.Thread(target=poll, args=(period, what)); th.start()
    time.sleep(0.02)
    r = [launches() for _ in range(3)]
    stop = True; th.join()
    print("launch us, poll period", period, "what", what, ["%.1f" % a for a in r])
