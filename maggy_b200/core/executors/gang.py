"""Multi-GPU trial gangs (new capability, BASELINE config 4: "4 trials x 2 GPUs").

The reference runs exactly one process per trial. Here a trial can own ``gpus_per_trial`` GPUs: the
pool starts that many processes per worker slot; rank 0 (the leader) owns the control ring and the
others follow it through a gang-private torch.distributed group (NCCL on GPUs, gloo on CPU) that is
initialised ONCE per worker lifetime, so consecutive trials pay no rendezvous cost. Inside train_fn
``torch.distributed`` is ready to use (e.g. wrap the model with maggy_b200.parallel.DataParallel).
"""
from __future__ import annotations

import datetime
import os


class Gang:
    def __init__(self, rank: int, size: int, port: int, timeout_s: int = 1800):
        import torch
        import torch.distributed as dist

        self.rank, self.size = rank, size
        local = int(os.environ.get("MAGGY_LOCAL_RANK", "0"))
        os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank),
                           "WORLD_SIZE": str(size), "LOCAL_RANK": str(local)})
        use_cuda = torch.cuda.is_available()
        if use_cuda:
            torch.cuda.set_device(local)  # the gang sees its own GPUs; member r drives cuda:r
        dist.init_process_group("nccl" if use_cuda else "gloo", rank=rank, world_size=size,
                                timeout=datetime.timedelta(seconds=timeout_s))
        self.dist = dist
        self.device = torch.device("cuda", local) if use_cuda else torch.device("cpu")

    @property
    def is_leader(self) -> bool:
        return self.rank == 0

    def broadcast(self, obj):
        """Leader -> followers object broadcast (trial assignment, stop sentinel)."""
        import cloudpickle

        # trial payloads may carry closures (ablation model/dataset generators): cloudpickle them ourselves,
        # torch's object collectives use the plain pickler
        box = [cloudpickle.dumps(obj) if self.rank == 0 else None]
        self.dist.broadcast_object_list(box, src=0, device=self.device)
        return cloudpickle.loads(box[0])

    def reset_agreement(self):
        self._pending = None

    def agree_stop(self, local_flag: bool) -> bool:
        """Collective OR of the members' stop flags with a one-call delay: returns the OR of the flags passed
        at the PREVIOUS call, so every member sees the same answer at the same call index, and the host never
        waits for more than the previous call's (long finished) all-reduce."""
        import torch

        prev = False
        pend = getattr(self, "_pending", None)
        if pend is not None:
            done, host = pend
            if done is not None:
                done.synchronize()
            prev = bool(int(host[0]) != 0)
        if self.device.type == "cuda":
            if not hasattr(self, "_side"):
                self._side = torch.cuda.Stream(device=self.device)
            flag = torch.full((1,), 1 if local_flag else 0, dtype=torch.int32, device=self.device)
            self.dist.all_reduce(flag, op=self.dist.ReduceOp.MAX)
            ready = torch.cuda.Event()
            ready.record()
            host = torch.zeros(1, dtype=torch.int32).pin_memory()
            with torch.cuda.stream(self._side):
                self._side.wait_event(ready)
                host.copy_(flag, non_blocking=True)
                done = torch.cuda.Event()
                done.record()
            flag.record_stream(self._side)
            self._pending = (done, host)
        else:
            flag = torch.tensor([1 if local_flag else 0], dtype=torch.int32)
            self.dist.all_reduce(flag, op=self.dist.ReduceOp.MAX)
            self._pending = (None, flag)
        return prev

    def mean(self, value: float) -> float:
        import torch

        t = torch.tensor([float(value)], dtype=torch.float64, device=self.device)
        if self.device.type == "cuda":
            t = t.float()
        self.dist.all_reduce(t)
        return float(t.item()) / self.size

    def close(self):
        try:
            if self.dist.is_initialized():
                self.dist.destroy_process_group()
        except Exception:
            pass
