"""Drop-in for the reference's ``truss2D_ENV`` module: ``ENV``, ``Game_research04`` (``_game_get_1_state``,
``_game_modify``, ``_set_model``, ``reset``, ``step``, ``re_game``) and ``pareto_state_data``
(reference ``test/*/code/truss2D_ENV.py:22-41, 207-604``).  Every evaluation is one batch-of-one call into
``libtfem.so``; the object model is refreshed from the results so ``savetxt`` and the driver's bookkeeping see
what the reference would have left behind."""
import random

import numpy as np

try:
    from set_seed_global import seedThis
    np.random.seed(seedThis)
    random.seed(seedThis)
except Exception:
    pass

from mop_truss_marl_b200.compat._backend import _np_or_py

MAX_MEM_NO = 5
MAX_FRONT = 50


def _degree_power(A, k):
    with np.errstate(divide="ignore"):
        deg = np.power(np.array(A.sum(1)), k).ravel()
    deg[np.isinf(deg)] = 0.0
    return np.diag(deg)


def pareto_state_data(pf, index=0):
    """chain graph over the current front (truss2D_ENV.py:22-41)"""
    n = len(pf)
    x_pf = np.zeros((n, 4), dtype=np.float32)
    for i in range(n):
        x_pf[i][0] = pf[i][0]
        x_pf[i][1] = pf[i][1]
        if i == index:
            x_pf[i][2] = 1
        x_pf[i][3] = n / MAX_FRONT
    A_pf = np.eye(n, dtype=np.float32)
    for i in range(n - 1):
        A_pf[i][i + 1] = 1
        A_pf[i + 1][i] = 1
    D_pf = _degree_power(A_pf, -1 / 2)
    return x_pf, np.matmul(D_pf, np.matmul(A_pf, D_pf))


class ENV:
    def __init__(self, game):
        self.name = "FRAME_ENV"
        self.game = game
        self.num_agents = game.num_agents
        self.over = 0
        self.output = []

    def check_over(self):
        if self.game.done_counter == 1:
            self.over = 1

    def reset(self):
        self.over = 0
        self.game.reset()
        self.output = []


class Game_research04:
    def __init__(self, end_step, model, num_agents=2):
        self.re_game(end_step, model, num_agents)

    def re_game(self, end_step, model, num_agents=2):
        self.name = "Game_research04"
        self.description = ("There are 2 type of agent \n Agent_s adjust node up and down\n"
                            " Agent_t adjust element section")
        self.objective = "min(Weigth),min(Diff_btw_targetShape_and_currentShape)"
        self.num_agents = num_agents
        self.gen_model = model
        self.num_x, self.num_y = model.num_x, model.num_y
        self.game_step = 1
        self.end_step = end_step
        self.height_change, self.topology_change = [], []
        self.max_y_val, self.min_y_val = model.y_max, model.y_min
        self.reward_counter = [0, 0]
        self.done_counter = 0
        self.current_hv = 0
        self.ref_point = [1, 1]
        self.front_max_distance = 0
        self.front_dis_distance = 0
        io = model._tfem.tab["int_obj"]                     # float32 sums of the generated geometry (:267-277)
        self.int_obj1, self.int_obj2 = np.float32(io[0]), np.float32(io[1])
        print("-------------------------------------------------------")
        print(self.description)
        print(self.objective)
        print("GAME WILL BE ENDED AFTER {} STEP".format(self.end_step))
        print("-------------------------------------------------------")

    # ---- helpers -------------------------------------------------------------------------------------
    def _constants(self):
        t = self.gen_model._tfem.tab
        return t["A_n"].copy(), t["mask"].copy(), t["nC_e"].copy()

    def _at_generated_geometry(self):
        m, t = self.gen_model.model, self.gen_model._tfem.tab
        last = len(self.gen_model.truss) - 1
        return (all(float(n.coord[1]) == t["y0"][i] and not isinstance(n.coord[1], np.floating)
                    for i, n in enumerate(m.nodes)) and all(e.section_no == last for e in m.elements))

    def _refresh_model(self, out, with_geometry=True):
        """write a batch-of-one libtfem result back into the object model"""
        gm, m = self.gen_model, self.gen_model.model
        if with_geometry:
            for i, n in enumerate(m.nodes):
                n.coord[1] = _np_or_py(float(out["y"][0, i]), bool(out["y_weak"][0, i]))
            for e, el in enumerate(m.elements):
                s = int(out["nN_x_e"][0, e, 0])
                el.section_no = s
                el.area = gm.truss[s][0] * 1e-4
                el.set_i(gm.truss[s][1] * 1e-8)
        gm.set_moveRange()
        gm._tfem.fill_results(m, out)

    # ---- the reference's entry points -------------------------------------------------------------------
    def _game_get_1_state(self, index=0):
        gm = self.gen_model
        A_n, mask, nC_e = self._constants()
        if self._at_generated_geometry():
            out = gm._tfem.reset_state()
            self._refresh_model(out, with_geometry=False)
        else:
            # current (non-generated) geometry: a neutral step keeps heights and sections
            N, E = len(gm.model.nodes), len(gm.model.elements)
            node_tab = np.zeros((N, 12), np.float32)
            node_tab[:, 1] = [n.coord[1] for n in gm.model.nodes]
            elem_tab = np.zeros((E, 21), np.float32)
            elem_tab[:, 0] = [e.section_no for e in gm.model.elements]
            a_geo = np.zeros((N, 2), np.float32)
            a_topo = np.tile(np.array([0, 0, 1], np.float32), (N, 1))
            mr = np.zeros((1, N, 2), np.float32)
            out = gm._tfem.game_modify(node_tab, elem_tab, mr, a_geo, a_topo, False)
            self._refresh_model(out)
        x_pf = np.zeros((1, 4), dtype=np.float32)
        x_pf[0] = [1, 1, 1, 1 / MAX_FRONT]
        A_pf = np.eye(1, dtype=np.float32)
        return (out["x_n"][0], A_n, out["A_s"][0], out["A_n_ts"][0], out["A_n_cs"][0], mask, x_pf, A_pf,
                out["nN_x_n"][0], out["nN_x_e"][0], nC_e)

    def _set_model(self, set_node, set_element):
        gm = self.gen_model
        for i, n in enumerate(gm.model.nodes):
            n.coord[1] = set_node[i][1]
        for i, el in enumerate(gm.model.elements):
            s = int(set_element[i][0])
            el.section_no = s
            el.area = gm.truss[s][0] * 1e-4
            el.set_i(gm.truss[s][1] * 1e-8)

    def _game_modify(self, set_node, set_element, nC_e, actions):
        gm = self.gen_model
        a_geo, a_topo = actions[0], actions[1]
        if not (isinstance(a_geo, np.ndarray) and a_geo.dtype == np.float32 and a_geo.flags["C_CONTIGUOUS"]
                and isinstance(a_topo, np.ndarray) and a_topo.dtype == np.float32 and a_topo.flags["C_CONTIGUOUS"]):
            raise TypeError("actions must be C-contiguous float32 arrays (what act() returns); they are clipped in place")
        # the move range the previous call left on the model (truss2D_ENV.py:405,410 read it before :557 resets it)
        mr = np.array([[[np.float32(n.max_up), np.float32(n.max_down)] for n in gm.model.nodes]], dtype=np.float32)
        coin = random.random() >= 0.5                           # :460
        out = gm._tfem.game_modify(set_node, set_element, mr, a_geo, a_topo, coin)
        self._refresh_model(out)
        A_n, mask, nC = self._constants()
        St_S = [out["x_n"][0], A_n, out["A_s"][0], out["A_n_ts"][0], out["A_n_cs"][0], mask, None, None,
                out["nN_x_n"][0], out["nN_x_e"][0], nC]
        p = out["point"][0]
        return [np.float32(p[0]), np.float32(p[1]), np.float32(p[2]), np.float32(p[3])], St_S

    def reset(self):
        self.height_change, self.topology_change = [], []
        self.reward_counter = [0, 0]
        self.done_counter = 0
        self.current_hv = 0
        self.ref_point = [1, 1]
        self.front_max_distance = 0
        self.front_dis_distance = 0

    def step(self):
        self.game_step += 1
