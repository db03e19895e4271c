"""PyTorch-level DDPG learner of the three node agents: the consumer on the far side of the hot path
(BASELINE.json configs[4]; SURVEY.md section 2 keeps it "PyTorch-level", only the gradient all-reduce is named).

Restates ``train/code/truss2D_RL.py``:
  * ``multimodes_actor``  (:49-127)  -> :class:`ActorNet`   (same parameters / names as the checkpoint and as
    ``tactor_create``; the rollout uses the CUDA kernel, this module is the differentiable twin for the update)
  * ``multimodes_critic`` (:130-266) -> :class:`CriticNet`  (10 + 11 GCNConv, sum-pool, concat, 3 Dense)
  * ``MADDPG.remember / train / update`` (:458-700) -> :class:`MADDPGLearner`

``GCNConv`` = ``A . (X . W) + b`` (spektral 1.2.0), ``GlobalSumPool`` = sum over nodes.  Quirks kept on purpose:
the Pareto embedding is tiled and ``tf.reshape``-d, not transposed (:89-95, :190-195); the critic target averages the
three per-agent next states (:611-618) and ALWAYS bootstraps (the terminal test at :606 compares an action array with
``1``); the actor step builds a NEW Adam every call (:625, so every step is Adam's
first step) with ``lr * 0.1`` and ``clipnorm=1``; targets move by ``tau = 0.005`` only every 1000 calls (:392-398).

When the replay / learner is partitioned over ranks, gradients are averaged with ONE flat-buffer all-reduce per model
update (:func:`allreduce_flat`): the only collective in the system (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import random
from collections import deque

import numpy as np
import torch
import torch.distributed as dist
from torch import nn

from .tf_checkpoint import ACTOR_LAYERS

ACTOR_SHAPES = {"gcn_l1_1": 13, "gcn_l1_2": 13, "gcn_l1_3": 13, "gcn_l1_4": 4}
CRITIC_L1_IN = (13, 13, 13, 4, 2, 3, 2, 3, 2, 3)          # gcn_l1_1 .. gcn_l1_10 (:176-213)


def _glorot_normal(fan_in, fan_out, gen):
    std = float(np.sqrt(2.0 / (fan_in + fan_out)))
    return torch.randn(fan_in, fan_out, generator=gen) * std


class _GCN(nn.Module):
    def __init__(self, fan_in, fan_out, gen):
        super().__init__()
        self.kernel = nn.Parameter(_glorot_normal(fan_in, fan_out, gen))
        self.bias = nn.Parameter(torch.zeros(fan_out))

    def forward(self, x, a):
        return torch.matmul(a, torch.matmul(x, self.kernel)) + self.bias


def _tile_reshape(pooled, n):
    """tf.ragged.stack([pooled]*N, axis=-1) -> [B,H,N]; tf.reshape(..., (B,N,H)) (not a transpose)"""
    b, h = pooled.shape
    return pooled.unsqueeze(-1).expand(b, h, n).reshape(b, n, h)


class ActorNet(nn.Module):
    """``multimodes_actor`` (truss2D_RL.py:49-127).  ``forward`` returns (geo [B,N,2], topo [B,N,3])."""

    def __init__(self, hidden=200, n_geo=2, n_topo=3, seed=20):
        super().__init__()
        gen = torch.Generator().manual_seed(seed)
        fan = lambda name: ACTOR_SHAPES.get(name, hidden)  # noqa: E731
        out = {"gcn_l4_1": n_geo, "gcn_l4_2": n_topo}
        self.layers = nn.ModuleDict({name: _GCN(fan(name), out.get(name, hidden), gen) for name in ACTOR_LAYERS})

    def forward(self, x_n, A_n, A_s, A_n_ts, A_n_cs, x_p, A_p):
        L, relu = self.layers, torch.relu
        if A_n.dim() == 2:
            A_n = A_n.unsqueeze(0).expand(x_n.shape[0], -1, -1)
        x11, x12, x13 = (relu(L[k](x_n, A_n)) for k in ("gcn_l1_1", "gcn_l1_2", "gcn_l1_3"))
        x14 = _tile_reshape(relu(L["gcn_l1_4"](x_p, A_p)).sum(dim=1), x_n.shape[1])
        s = (relu(L["gcn_l2_1"](x11, A_n)) + relu(L["gcn_l2_2"](x12, A_n_ts)) + relu(L["gcn_l2_3"](x12, A_n_cs))
             + relu(L["gcn_l2_4"](x13, A_s)) + relu(L["gcn_l2_5"](x14, A_n)))
        x31, x32 = relu(L["gcn_l3_1"](s, A_n)), relu(L["gcn_l3_2"](s, A_s))
        return torch.sigmoid(L["gcn_l4_1"](x31, A_n)), torch.sigmoid(L["gcn_l4_2"](x32, A_n))

    # ---- exchange with the CUDA actor / the TF checkpoint: {layer: (kernel [in,out], bias [out])} float32 numpy ----
    def export_weights(self):
        return {k: (m.kernel.detach().cpu().numpy().astype(np.float32).copy(),
                    m.bias.detach().cpu().numpy().astype(np.float32).copy()) for k, m in self.layers.items()}

    def import_weights(self, weights):
        with torch.no_grad():
            for k, m in self.layers.items():
                m.kernel.copy_(torch.as_tensor(np.asarray(weights[k][0]), dtype=m.kernel.dtype))
                m.bias.copy_(torch.as_tensor(np.asarray(weights[k][1]), dtype=m.bias.dtype))


class CriticNet(nn.Module):
    """``multimodes_critic`` (truss2D_RL.py:130-266): Q(state, own action, the two other agents' actions) [B,1]."""

    def __init__(self, hidden=200, n_q=200, seed=21):
        super().__init__()
        gen = torch.Generator().manual_seed(seed)
        self.l1 = nn.ModuleList([_GCN(f, hidden, gen) for f in CRITIC_L1_IN])
        self.l2 = nn.ModuleList([_GCN(hidden, hidden, gen) for _ in range(11)])
        self.dense_1, self.dense_2, self.dense_out = nn.Linear(11 * hidden, n_q), nn.Linear(n_q, n_q), nn.Linear(n_q, 1)
        with torch.no_grad():
            for d in (self.dense_1, self.dense_2):
                d.weight.copy_(_glorot_normal(d.in_features, d.out_features, gen).t())
                d.bias.zero_()
            lim = float(np.sqrt(6.0 / (n_q + 1)))            # Dense(1): keras default glorot_uniform
            self.dense_out.weight.copy_((torch.rand(1, n_q, generator=gen) * 2 - 1) * lim)
            self.dense_out.bias.zero_()

    def forward(self, x_n, A_n, A_s, A_n_ts, A_n_cs, x_p, A_p, self_g, self_t, o1_g, o1_t, o2_g, o2_t):
        relu = torch.relu
        if A_n.dim() == 2:
            A_n = A_n.unsqueeze(0).expand(x_n.shape[0], -1, -1)
        n = x_n.shape[1]
        x1 = [relu(self.l1[i](x_n, A_n)) for i in range(3)]
        x14 = _tile_reshape(relu(self.l1[3](x_p, A_p)).sum(dim=1), n)
        acts = [relu(self.l1[4 + i](a, A_n)) for i, a in enumerate((self_g, self_t, o1_g, o1_t, o2_g, o2_t))]
        x2 = [relu(self.l2[0](x1[0], A_n)), relu(self.l2[1](x1[1], A_n_ts)), relu(self.l2[2](x1[1], A_n_cs)),
              relu(self.l2[3](x1[2], A_s))]
        x2 += [relu(self.l2[4 + i](a, A_n)) for i, a in enumerate(acts)]
        x2.append(relu(self.l2[10](x14, A_n)))
        q = torch.cat([t.sum(dim=1) for t in x2], dim=-1)     # GlobalSumPool + Concatenate -> [B, 11*hidden]
        return self.dense_out(relu(self.dense_2(relu(self.dense_1(q)))))


def allreduce_flat(params, group=None):
    """Average the gradients of ``params`` over the process group with ONE all-reduce on a flat buffer."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 0
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return 0
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat /= dist.get_world_size(group)
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()
    return flat.numel()


def _clip_global_norm(params, max_norm=1.0):
    """keras ``clipnorm``: every gradient tensor is clipped to ``max_norm`` on its OWN norm (branch-free on the device: no
    host synchronisation, capturable in a CUDA graph)"""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    norms = torch._foreach_norm(grads)
    scales = [torch.clamp(max_norm / (n + 1e-30), max=1.0) for n in norms]
    torch._foreach_mul_(grads, scales)


class _Agent:
    def __init__(self, lr, hidden, n_q, seed, device):
        self.actor, self.target_actor = ActorNet(hidden, seed=seed).to(device), ActorNet(hidden, seed=seed).to(device)
        self.critic, self.target_critic = CriticNet(hidden, n_q, seed=seed + 1).to(device), CriticNet(hidden, n_q, seed=seed + 1).to(device)
        self.lr, self.tau, self.update_num = lr, 0.005, 0
        self.critic_opt = torch.optim.Adam(self.critic.parameters(), lr=lr, eps=1e-7,
                                           capturable=torch.device(device).type == "cuda")
        self.update_init()

    @staticmethod
    def _blend(dst, src, tau):
        with torch.no_grad():
            for d, s in zip(dst.parameters(), src.parameters()):
                d.copy_(s if tau >= 1.0 else s * tau + d * (1.0 - tau))

    def update_init(self):
        self._blend(self.target_actor, self.actor, 1.0)
        self._blend(self.target_critic, self.critic, 1.0)

    def update(self):
        """``multimodals_OneAgent.update`` (:392-400): soft update only when ``update_num == 1000``"""
        if self.update_num == 1000:
            self._blend(self.target_actor, self.actor, self.tau)
            self._blend(self.target_critic, self.critic, self.tau)
            self.update_num = 0
            return True
        return False


STATE_KEYS = ("x_n", "A_n", "A_s", "A_n_ts", "A_n_cs", "x_p", "A_p")


class DeviceReplay:
    """Replay memory resident on the learner's device: a ring of ``capacity`` transitions in preallocated tensors, filled
    from whole environment batches without a host round trip (``push``) and sampled by index (``sample``).  Same content
    as ``MADDPG.remember`` rows (``train/code/truss2D_RL.py:439-456``): state, the three agents' actions, the three rewards,
    the three next states, done.  ``A_n`` is a topology constant and kept once."""

    def __init__(self, capacity, N, P, device, seed=0):
        self.capacity, self.N, self.P, self.device = int(capacity), int(N), int(P), torch.device(device)
        f = dict(dtype=torch.float32, device=self.device)
        C_ = self.capacity

        def state():
            return {"x_n": torch.zeros(C_, N, 13, **f), "A_s": torch.zeros(C_, N, N, **f), "A_n_ts": torch.zeros(C_, N, N, **f),
                    "A_n_cs": torch.zeros(C_, N, N, **f), "x_p": torch.zeros(C_, P, 4, **f), "A_p": torch.zeros(C_, P, P, **f)}
        self.state, self.next_states = state(), [state() for _ in range(3)]
        self.geo = [torch.zeros(C_, N, 2, **f) for _ in range(3)]
        self.topo = [torch.zeros(C_, N, 3, **f) for _ in range(3)]
        self.rewards, self.done = torch.zeros(C_, 3, **f), torch.zeros(C_, **f)
        self.A_n = None
        self.size, self.head = 0, 0
        self.gen = torch.Generator(device=self.device).manual_seed(seed)

    def __len__(self):
        return self.size

    def push(self, state, actions, rewards, next_states, done, rows=None):
        """``state`` / ``next_states[k]``: dicts (or 7-tuples in STATE_KEYS order) of device tensors with a leading batch
        axis; ``rows``: optional index tensor selecting which environments of the batch are stored"""
        def as_dict(s):
            return s if isinstance(s, dict) else dict(zip(STATE_KEYS, s))
        state, next_states = as_dict(state), [as_dict(s) for s in next_states]
        if self.A_n is None:
            self.A_n = state["A_n"].detach().clone()
        pick = (lambda t: t) if rows is None else (lambda t: t.index_select(0, rows))
        n = int(rewards.shape[0] if rows is None else rows.numel())
        if n > self.capacity:
            raise ValueError("more transitions than the replay holds")
        idx = (torch.arange(n, device=self.device) + self.head) % self.capacity
        for k in self.state:
            self.state[k].index_copy_(0, idx, pick(state[k]))
            for j in range(3):
                self.next_states[j][k].index_copy_(0, idx, pick(next_states[j][k]))
        for j in range(3):
            self.geo[j].index_copy_(0, idx, pick(actions[j][0]))
            self.topo[j].index_copy_(0, idx, pick(actions[j][1]))
        self.rewards.index_copy_(0, idx, pick(rewards))
        d = done if torch.is_tensor(done) else torch.full((rewards.shape[0],), float(done), device=self.device)
        self.done.index_copy_(0, idx, pick(d.to(torch.float32)))
        self.head = (self.head + n) % self.capacity
        self.size = min(self.capacity, self.size + n)

    def draw(self, batch_size):
        """indices of a uniform sample of the stored transitions"""
        return torch.randint(0, self.size, (batch_size,), device=self.device, generator=self.gen)

    def sample(self, batch_size, idx=None):
        if idx is None:
            idx = self.draw(batch_size)

        def st(d):
            return (d["x_n"][idx], self.A_n, d["A_s"][idx], d["A_n_ts"][idx], d["A_n_cs"][idx], d["x_p"][idx], d["A_p"][idx])
        return (st(self.state), [st(d) for d in self.next_states], [(self.geo[j][idx], self.topo[j][idx]) for j in range(3)],
                self.rewards[idx], self.done[idx])


class MADDPGLearner:
    """Replay + update of the three agents (``MADDPG.remember / train / update``).

    A transition is ``(state, [(geo_k, topo_k)]*3, rewards[3], [next_state_k]*3, done)`` where a state is the tuple
    ``(x_n, A_n, A_s, A_n_ts, A_n_cs, x_p, A_p)`` of one environment; ``remember_batch`` takes the same with a leading
    batch axis (what the batched environment produces).  ``train()`` samples ``batch_size`` transitions and does, per
    agent, the critic regression and the actor ascent; gradients are all-reduced when a process group is initialised.
    """

    def __init__(self, lr=1e-7, gamma=0.99, hidden=200, n_q=200, max_mem=100000, batch_size=32, device="cpu", seed=20):
        self.device = torch.device(device)
        self.gamma, self.batch_size = gamma, batch_size
        self.agents = [_Agent(lr, hidden, n_q, seed + 10 * k, self.device) for k in range(3)]
        self.memory = deque(maxlen=max_mem)
        self.rng = random.Random(seed)
        self.allreduced_elements = 0
        self.last_losses = None
        self.honour_done = False           # False = the reference's behaviour (its terminal test never fires, see train())
        self._graph, self._graph_idx = None, None
        self.device_replay = None          # a DeviceReplay: train() samples from it instead of the host-side deque
        self.keep_losses_on_device = False  # True: no host synchronisation inside train() (last_losses holds tensors)

    # ------------------------------------------------------------------------------------------------
    def _to(self, a):
        return torch.as_tensor(np.asarray(a), dtype=torch.float32)

    def remember(self, state, actions, rewards, next_states, done):
        st = tuple(self._to(s) for s in state)
        self.memory.append((st, tuple((self._to(g), self._to(t)) for g, t in actions), self._to(rewards),
                            tuple(tuple(self._to(s) for s in ns) for ns in next_states), float(done)))

    def remember_batch(self, state, actions, rewards, next_states, done):
        """the same with a leading batch axis on every array (``A_n`` may stay [N,N]): what the batched environment
        and actor produce; ``rewards`` [B,3], ``done`` [B] or a scalar"""
        rewards = np.asarray(rewards)

        def row(tup, i):
            return tuple(np.asarray(s) if (k == 1 and np.asarray(s).ndim == 2) else np.asarray(s)[i] for k, s in enumerate(tup))
        for i in range(rewards.shape[0]):
            self.remember(row(state, i), [(np.asarray(g)[i], np.asarray(t)[i]) for g, t in actions], rewards[i],
                          [row(ns, i) for ns in next_states], np.asarray(done).reshape(-1)[i] if np.ndim(done) else done)

    # ------------------------------------------------------------------------------------------------
    def _stack_state(self, states):
        return tuple(torch.stack([s[j] for s in states]).to(self.device) for j in range(len(STATE_KEYS)))

    def train(self):
        """one ``MADDPG.train()`` call (:463-689); returns False while the replay holds fewer than ``batch_size``"""
        if self.device_replay is not None:
            if len(self.device_replay) < self.batch_size:
                return False
            if self._graph is not None:                      # the whole update as ONE captured CUDA graph
                self._graph_idx.copy_(self.device_replay.draw(self.batch_size))
                self._graph.replay()
                return True
            S, NS, A, R, done = self.device_replay.sample(self.batch_size)
        else:
            if len(self.memory) < self.batch_size:
                return False
            samples = self.rng.sample(list(self.memory), self.batch_size)
            S = self._stack_state([m[0] for m in samples])
            NS = [self._stack_state([m[3][k] for m in samples]) for k in range(3)]
            A = [(torch.stack([m[1][k][0] for m in samples]).to(self.device),
                  torch.stack([m[1][k][1] for m in samples]).to(self.device)) for k in range(3)]
            R = torch.stack([m[2] for m in samples]).to(self.device)                    # [B,3]
            done = torch.tensor([m[4] for m in samples], device=self.device)
        self._update_from_batch(S, NS, A, R, done)
        return True

    def capture_graph(self):
        """Capture one whole ``train()`` -- sampling by index, the nine target-actor and nine target-critic forwards, three
        critic regressions and three actor ascents with their gradient all-reduces and optimiser steps (~6 000 small kernels
        at batch 32) -- into a CUDA graph; later ``train()`` calls draw new indices and replay it.  Needs the device
        replay and a CUDA device.  The three warm-up updates PyTorch asks for before a capture run on a side stream and
        are undone afterwards (models and optimiser states are restored), so capturing does not change the training."""
        if self.device_replay is None or self.device.type != "cuda" or len(self.device_replay) < self.batch_size:
            raise RuntimeError("capture_graph needs a filled DeviceReplay on a CUDA device")
        self.keep_losses_on_device = True                   # no host read-back inside a capture
        models = [m for a in self.agents for m in (a.actor, a.critic, a.target_actor, a.target_critic)]
        saved = [{k: v.clone() for k, v in m.state_dict().items()} for m in models]
        saved_opt = [{k: (v.clone() if torch.is_tensor(v) else v) for k, v in st.items()}
                     for a in self.agents for st in a.critic_opt.state.values()]
        self._graph_idx = self.device_replay.draw(self.batch_size)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(3):
                self._update_from_batch(*self.device_replay.sample(self.batch_size, idx=self._graph_idx))
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self._update_from_batch(*self.device_replay.sample(self.batch_size, idx=self._graph_idx))
        with torch.no_grad():                               # undo the warm-up and capture-time updates
            for m, sd in zip(models, saved):
                m.load_state_dict(sd)
            it = iter(saved_opt)
            for a in self.agents:
                for st in a.critic_opt.state.values():
                    old = next(it, None)
                    for k, v in st.items():
                        if torch.is_tensor(v):
                            if old is not None and k in old:
                                v.copy_(old[k])
                            else:
                                v.zero_()
            if not saved_opt:                               # the optimiser had no state yet: back to step 0
                for a in self.agents:
                    for st in a.critic_opt.state.values():
                        for v in st.values():
                            if torch.is_tensor(v):
                                v.zero_()
        self._graph = graph
        return graph

    def _update_from_batch(self, S, NS, A, R, done):
        order = {0: (0, 1, 2), 1: (1, 0, 2), 2: (2, 0, 1)}                           # own action first (:560-562)
        with torch.no_grad():
            # next actions of every target actor on every agent's next state, then the target critics (:564-606)
            nxt = [[self.agents[j].target_actor(*NS[f]) for j in range(3)] for f in range(3)]
            q_next = []
            for k in range(3):
                o = order[k]
                qs = [self.agents[k].target_critic(*NS[f], *nxt[f][o[0]], *nxt[f][o[1]], *nxt[f][o[2]]) for f in range(3)]
                q_next.append((qs[0] + qs[1] + qs[2]) / 3.0)
        losses = []
        for k, ag in enumerate(self.agents):
            o = order[k]
            # the reference's terminal test reads row element 4 of the replay row -- an action array, never the int 1 -- so
            # `done is 1` is always False and EVERY target bootstraps (:606-613); kept (``done`` is stored but not used),
            # set ``self.honour_done = True`` for the textbook target
            y = R[:, k:k + 1] + self.gamma * q_next[k]
            if self.honour_done:
                y = torch.where(done.view(-1, 1) == 1.0, R[:, k:k + 1], y)
            # critic: train_on_batch with mse, Adam(lr, clipnorm=1) (:619)
            ag.critic_opt.zero_grad(set_to_none=True)
            q = ag.critic(*S, *A[o[0]], *A[o[1]], *A[o[2]])
            c_loss = torch.mean((q - y) ** 2)
            c_loss.backward()
            self.allreduced_elements += allreduce_flat(list(ag.critic.parameters()))
            _clip_global_norm(ag.critic.parameters(), 1.0)
            ag.critic_opt.step()
            # actor: maximise the critic's value of the three CURRENT actors' actions; gradient to agent k only (:620-625)
            for p in ag.actor.parameters():
                p.grad = None
            pred = [self.agents[j].actor(*S) for j in range(3)]
            a_loss = -ag.critic(*S, *pred[o[0]], *pred[o[1]], *pred[o[2]]).mean()
            grads = torch.autograd.grad(a_loss, list(ag.actor.parameters()), allow_unused=True)
            for p, g in zip(ag.actor.parameters(), grads):
                p.grad = g if g is not None else torch.zeros_like(p)
            self.allreduced_elements += allreduce_flat(list(ag.actor.parameters()))
            _clip_global_norm(ag.actor.parameters(), 1.0)
            # a NEW Adam every call (:625): Keras applies lr sqrt(1 - b2) / (1 - b1) * m / (sqrt(v) + eps) with m = (1 - b1) g,
            # v = (1 - b2) g^2 on the first step, i.e. lr * g / (|g| + eps / sqrt(1 - b2)), eps = 1e-7, b2 = 0.999
            with torch.no_grad():
                params = list(ag.actor.parameters())
                grads = [p.grad for p in params]
                denom = torch._foreach_abs(grads)
                torch._foreach_add_(denom, 1e-7 / (1.0 - 0.999) ** 0.5)
                torch._foreach_addcdiv_(params, grads, denom, value=-(ag.lr * 0.1))
            losses.append((c_loss.detach(), a_loss.detach()) if self.keep_losses_on_device
                          else (float(c_loss.detach()), float(a_loss.detach())))
        self.last_losses = losses

    def update(self):
        """``MADDPG.update`` (:692-697): every 300 actor calls let each agent check its own counter"""
        interval = len(self.agents) * 100
        if self.agents[0].update_num % interval == 0:
            return [a.update() for a in self.agents]
        return None

    def actor_tensors(self, k):
        """agent k's online actor as ``{layer: (kernel, bias)}`` of its parameter tensors (no copy): for
        ``BatchedActor.set_weights_device`` when the learner lives on the actor's GPU"""
        return {name: (m.kernel.data, m.bias.data) for name, m in self.agents[k].actor.layers.items()}

    def actor_weights(self, k):
        """weights of agent k's online actor -- the model ``act`` uses (:340) -- for ``BatchedActor.set_weights``"""
        return self.agents[k].actor.export_weights()
