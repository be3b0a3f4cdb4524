import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from maggy_b200 import Trial
from maggy_b200.core import rpc
from maggy_b200.core.sched_kernel import DeviceScheduler

d = rpc.DriverChannel(2, ring_slots=64, pump=False)
d.configure_policy(False, "max", 0, 1)
sched = DeviceScheduler(d, direction="max", es_min=0, es_interval=1, max_steps=32, cap_final=64)
sched.start()
host, h = d.lib, d.h
print("offsets ctl", host.mgr_off_ctl(h, 0), "host0", host.mgr_off_host_ring(h, 0), "host1", host.mgr_off_host_ring(h, 1),
      "dev0", host.mgr_off_dev_ring(h, 0), "dev1", host.mgr_off_dev_ring(h, 1), "stride", host.mgr_ring_stride(h), flush=True)
w0, w1 = rpc.WorkerChannel(d.name, 0), rpc.WorkerChannel(d.name, 1)
time.sleep(0.1); print("start", sched.stats(), flush=True)
good = Trial({"g": 1})
for s in range(4):
    w0.post_metric(good.trial_id, 0.9, s)
    time.sleep(0.05); print("after metric", s, sched.stats(), flush=True)
w0.post_final(good.trial_id, 0.9)
time.sleep(0.1); print("after final", sched.stats(), flush=True)
base = host.mgr_base(h)
recs = (rpc.MgRecord * 8).from_address(base + host.mgr_off_host_ring(h, 0) + 128)
print("host ring w0:", [(r.seq, r.type, r.step) for r in recs][:7], flush=True)
bad = Trial({"b": 1})
w1.post_metric_device(bad.trial_id, torch.tensor(0.1, device="cuda"), 0)
time.sleep(0.3); print("after dev metric", sched.stats(), "stop?", w1.check_stop(bad.trial_id), flush=True)
recs = (rpc.MgRecord * 4).from_address(base + host.mgr_off_dev_ring(h, 1) + 128)
print("dev ring w1:", [(r.seq, r.type, r.step, r.value) for r in recs][:3], flush=True)
print("final", sched.stop())
