"""Host-resident rollouts: the caller keeps the reference's state tuple in (pinned) host memory, like the reference
driver does (``master_DDPG_truss2D_MO.py:167-260``), and every step moves it to the GPU, lets the actor act, steps
the environments and brings the new state tuple, the objective point and the actions back.

The environments are independent, so the batch is cut into pieces and the three legs -- host->device copies, the
two C-ABI calls (``tactor_act``, ``tfem_step``) and device->host copies -- run on three CUDA streams: piece i+1 is
uploading while piece i computes and piece i-1 downloads (PCIe is full duplex).  PyTorch provides the streams,
events and pinned buffers only.
"""
from __future__ import annotations

import torch

STATE_IN = ("x_n", "A_s", "A_n_ts", "A_n_cs", "nN_x_n", "nN_x_e", "move_range")
STATE_OUT = ("x_n", "A_s", "A_n_ts", "A_n_cs", "nN_x_n", "nN_x_e", "move_range", "point", "status")


class HostRollout:
    """``env``: BatchedTrussEnv, ``actor``: BatchedActor of the same batch; ``pieces``: how many pieces the batch
    is cut into (each should still fill the GPU: >= 148 tiles of 128 rows, i.e. >= ~1184 small / 592 large envs)."""

    def __init__(self, env, actor, pieces: int = 2):
        self.env, self.actor = env, actor
        B, dev = env.B, env.device
        pieces = max(1, min(int(pieces), B))
        step = -(-B // pieces)
        step = -(-step // 32) * 32                 # piece boundaries on 32 environments: every sub-array stays 16-byte aligned
        self.ranges = [(lo, min(lo + step, B)) for lo in range(0, B, step)]
        self.s_in, self.s_run, self.s_out = (torch.cuda.Stream(device=dev) for _ in range(3))
        self.ev_in = [torch.cuda.Event() for _ in self.ranges]
        self.ev_run = [torch.cuda.Event() for _ in self.ranges]
        self.a_geo = torch.empty(B, env.N, 2, device=dev)
        self.a_topo = torch.empty(B, env.N, 3, device=dev)
        self.coin = torch.empty(B, dtype=torch.uint8, device=dev)

    def alloc_host(self):
        """pinned host buffers for one state tuple (+ point, status, actions)"""
        env = self.env
        h = {k: torch.empty(getattr(env, k).shape, dtype=getattr(env, k).dtype).pin_memory() for k in STATE_OUT}
        h["a_geo"] = torch.empty(env.B, env.N, 2).pin_memory()
        h["a_topo"] = torch.empty(env.B, env.N, 3).pin_memory()
        return h

    def bytes_per_step(self):
        env = self.env
        nbytes = lambda t: t.numel() * t.element_size()  # noqa: E731
        h2d = sum(nbytes(getattr(env, k)) for k in STATE_IN) + env.B
        d2h = sum(nbytes(getattr(env, k)) for k in STATE_OUT) + nbytes(self.a_geo) + nbytes(self.a_topo)
        return h2d, d2h

    def step(self, state_host: dict, coin_host: torch.Tensor, x_p: torch.Tensor, A_p: torch.Tensor, out_host: dict,
             n_pf=None):
        """state_host[k] for k in STATE_IN and coin_host [B] uint8: pinned host tensors; x_p/A_p/n_pf: the Pareto-front
        graph (device, [B,P,4] / [B,P,P] / [B]).  Fills out_host (STATE_OUT + a_geo, a_topo) and returns it after the
        last piece has landed."""
        env, actor = self.env, self.actor
        cur = torch.cuda.current_stream(env.device)
        for s in (self.s_in, self.s_run, self.s_out):
            s.wait_stream(cur)
        for i, (lo, hi) in enumerate(self.ranges):
            with torch.cuda.stream(self.s_in):
                for k in STATE_IN:
                    getattr(env, k)[lo:hi].copy_(state_host[k][lo:hi], non_blocking=True)
                self.coin[lo:hi].copy_(coin_host[lo:hi], non_blocking=True)
                self.ev_in[i].record(self.s_in)
            with torch.cuda.stream(self.s_run):
                self.s_run.wait_event(self.ev_in[i])
                actor.act(env.x_n[lo:hi], env.A_n, env.A_s[lo:hi], env.A_n_ts[lo:hi], env.A_n_cs[lo:hi],
                          x_p[lo:hi], A_p[lo:hi], n_pf=None if n_pf is None else n_pf[lo:hi],
                          out=(self.a_geo[lo:hi], self.a_topo[lo:hi]))
                env.step(self.a_geo[lo:hi], self.a_topo[lo:hi], self.coin[lo:hi], rows=(lo, hi))
                self.ev_run[i].record(self.s_run)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(self.ev_run[i])
                for k in STATE_OUT:
                    out_host[k][lo:hi].copy_(getattr(env, k)[lo:hi], non_blocking=True)
                out_host["a_geo"][lo:hi].copy_(self.a_geo[lo:hi], non_blocking=True)
                out_host["a_topo"][lo:hi].copy_(self.a_topo[lo:hi], non_blocking=True)
        self.s_out.synchronize()
        cur.wait_stream(self.s_run)
        return out_host
