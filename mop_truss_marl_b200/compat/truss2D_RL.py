"""Drop-in for the actor side of the reference's ``truss2D_RL`` module: ``MADDPG(...)``,
``.agents[k].act(...)``, ``.agents[k].actor_model.load_weights(prefix).expect_partial()``
(reference ``train/code/truss2D_RL.py:268-462``).  The forward pass runs on the GPU (``tactor_forward``,
batch of one); the Ornstein-Uhlenbeck noise is drawn on the host with ``np.random.randn(1)`` per entry in the
reference's order, so a seeded run consumes the NumPy stream exactly like the reference.

The critic, the replay sampling and ``train`` are out of scope for this hot path:
``remember`` stores, ``train`` raises."""
from collections import deque

import numpy as np

from mop_truss_marl_b200 import tf_checkpoint

try:
    from set_seed_global import seedThis
    np.random.seed(seedThis)
except Exception:
    seedThis = 20


class OUNoise:
    def __init__(self, mu, theta, sigma):
        self.mu, self.theta, self.sigma, self.dt = mu, theta, sigma, 0.0001

    def gen_noise(self, x):
        return self.theta * (self.mu - x) * self.dt + self.sigma * np.random.randn(1)


class _LoadStatus:
    def expect_partial(self):
        return self


class _ActorModel:
    """stands in for the Keras model: holds the 13 GCN layers' weights and the device actor"""

    def __init__(self, hidden):
        self.weights = tf_checkpoint.random_actor_weights(seed=seedThis, hidden=hidden)
        self._device_actor = {}

    def load_weights(self, prefix):
        self.weights = tf_checkpoint.load_actor_weights(prefix)
        self._device_actor = {}
        return _LoadStatus()

    def device_actor(self, nodes):
        if nodes not in self._device_actor:
            from mop_truss_marl_b200.actor import BatchedActor
            self._device_actor[nodes] = BatchedActor(self.weights, nodes, max_batch=1, sigma=0.0, theta=0.0)
        return self._device_actor[nodes]


class _CriticModel:
    def load_weights(self, prefix):
        return _LoadStatus()        # critic is not part of the accelerated path


class multimodals_OneAgent:
    def __init__(self, lr, ep, epd, gamma, a_nn, c_nn, num_action1, num_action2, mu_s, theta_s, sigma_s,
                 mu_t, theta_t, sigma_t, all_agent, batch):
        self.number = 1
        self.lr, self.gamma, self.a_nn, self.c_nn = lr, gamma, a_nn, c_nn
        self.num_action1, self.num_action2 = num_action1, num_action2
        self.batch_size = batch
        self.update_num = 0
        self.noise_geo = [OUNoise(mu_s[i], theta_s[i], sigma_s[i]) for i in range(num_action1)]
        self.noise_topo = [OUNoise(mu_t[i], theta_t[i], sigma_t[i]) for i in range(num_action2)]
        self.actor_model = _ActorModel(a_nn)
        self.target_actor_model = _ActorModel(a_nn)
        self.critic_model = _CriticModel()
        self.target_critic_model = _CriticModel()
        self.all_agent = all_agent

    def act(self, x_n, A_n, A_s, A_n_ts, A_n_cs, x_p, A_p):
        import torch
        act = self.actor_model.device_actor(int(np.asarray(x_n).shape[0]))
        dev = act.device
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)   # noqa: E731
        geo, topo = act.forward(t(x_n)[None].contiguous(), t(A_n), t(A_s)[None].contiguous(),
                                t(A_n_ts)[None].contiguous(), t(A_n_cs)[None].contiguous(),
                                t(x_p)[None].contiguous(), t(A_p)[None].contiguous())
        action_geo = geo[0].cpu().numpy()
        action_topo = topo[0].cpu().numpy()
        if self.noise_geo is not None:
            for i in range(len(action_geo)):
                for j in range(len(action_geo[i])):
                    action_geo[i][j] += self.noise_geo[j].gen_noise(action_geo[i][j])[0]
        if self.noise_topo is not None:
            for i in range(len(action_topo)):
                for j in range(len(action_topo[i])):
                    action_topo[i][j] += self.noise_topo[j].gen_noise(action_topo[i][j])[0]
        self.update_num += 1
        return action_geo, action_topo


class MADDPG:
    def __init__(self, lr, ep, epd, gamma, a_nn, c_nn, max_mem, num_agents, num_action, mu, theta, sigma,
                 max_poss_n_num=1):
        self.num_agents = num_agents
        self.lr, self.epint, self.ep, self.epd, self.epmin, self.gamma = lr, ep, ep, epd, 0.05, gamma
        self.a_nn, self.c_nn = a_nn, c_nn
        self.mu, self.theta, self.sigma = mu, theta, sigma
        self.temprp = deque(maxlen=max_mem)
        for _ in range(max_poss_n_num):
            self.temprp.append(deque(maxlen=max_mem))
        self.num_state = [0, 0]
        self.num_action = num_action
        self.batch_size = 32
        self.max_poss_n_num = max_poss_n_num
        self.agents, self.update_counter = [], []
        for k in range(3):                                   # the reference always builds three (:423-441)
            a = multimodals_OneAgent(lr, ep, epd, gamma, a_nn, c_nn, num_action[0], num_action[1], mu[0], theta[0],
                                     sigma[0], mu[1], theta[1], sigma[1], num_agents, self.batch_size)
            a.number = k + 1
            self.agents.append(a)
            self.update_counter.append(0)

    def remember(self, state, a0_g, a0_t, a1_g, a1_t, a2_g, a2_t, reward, next_state1, next_state2, next_state3,
                 done, n_node):
        t = [state, a0_g, a0_t, a1_g, a1_t, a2_g, a2_t, reward, next_state1, next_state2, next_state3, done]
        self.temprp[0].append(t)
        self._pending = getattr(self, "_pending", [])
        self._pending.append(t)                              # handed to the learner's replay at the next train()

    # ---- learner: PyTorch-level restatement of MADDPG.train / update (mop_truss_marl_b200/learner.py) ----
    def _learner(self):
        if getattr(self, "_lrn", None) is None:
            import torch
            from mop_truss_marl_b200.learner import MADDPGLearner
            dev = "cuda" if torch.cuda.is_available() else "cpu"
            self._lrn = MADDPGLearner(lr=self.lr, gamma=self.gamma, hidden=self.a_nn, n_q=self.c_nn,
                                      max_mem=self.temprp[0].maxlen, batch_size=self.batch_size, device=dev, seed=seedThis)
            for k, a in enumerate(self.agents):              # carry over loaded checkpoints (critic weights do not exist)
                self._lrn.agents[k].actor.import_weights(a.actor_model.weights)
                self._lrn.agents[k].update_init()
        return self._lrn

    def train(self):
        lrn = self._learner()
        # move the transitions remembered since the last call into the learner's replay
        fresh, self._pending = getattr(self, "_pending", []), []
        for t in fresh:
            state = [t[0][i] for i in (0, 1, 2, 3, 4, 6, 7)]                      # the reference's tuple carries mask at [5]
            nxt = [[ns[i] for i in (0, 1, 2, 3, 4, 6, 7)] for ns in (t[8], t[9], t[10])]
            lrn.remember(state, [(t[1], t[2]), (t[3], t[4]), (t[5], t[6])], np.asarray(t[7], dtype=np.float32), nxt,
                         1 if t[11] == 1 else 0)
        if not lrn.train():
            return
        for k, a in enumerate(self.agents):                  # the rollout acts with the updated online actor (:340)
            a.actor_model.weights = lrn.actor_weights(k)
            for dev_actor in a.actor_model._device_actor.values():
                dev_actor.set_weights(a.actor_model.weights)

    def update(self):
        interval = len(self.agents) * 100                    # (:692-697)
        if self.agents[0].update_num % interval == 0 and getattr(self, "_lrn", None) is not None:
            for k, a in enumerate(self.agents):
                self._lrn.agents[k].update_num = a.update_num
                if self._lrn.agents[k].update():
                    a.update_num = 0
