"""Drop-in for the reference's ``truss2D_GEN`` module (``gen_model``, reference ``truss2D_GEN.py:42-434``).

The mesh, supports, loads and targets come from libtfem's family tables (``tfem_get_table``); the object graph
(``.model.nodes[i].coord/res/top_node/target/...``, ``.model.elements[i].area/length/section_no/...``) is rebuilt
from them so the driver and ``savetxt`` see the attributes they expect.  The star-import also provides ``os``,
``pd`` and ``plt`` because the reference driver uses them without importing them
(``master_DDPG_truss2D_MO.py:129-136``)."""
import csv
import os  # noqa: F401  (re-exported on purpose)
import random

import numpy as np

try:
    import pandas as pd  # noqa: F401
except Exception:        # pragma: no cover
    pd = None
try:
    import matplotlib.pyplot as plt  # noqa: F401
except Exception:
    plt = None

from FEM_2Dtruss import Element, Load, Model, Node
from mop_truss_marl_b200.compat import _backend
from mop_truss_marl_b200.families import SECTION_AREA_CM2, SECTION_INERTIA_CM4, FamilySpec

try:
    from set_seed_global import seedThis
    random.seed(seedThis)
    np.random.seed(seedThis)
except Exception:
    pass


def read_section(path):
    with open(path, newline="") as f:
        data = list(csv.reader(f))
    return np.array(data).astype(float)


class gen_model:
    SYMMETRY = None          # set to families.SYM_* to force a convention; default: by num_x

    def __init__(self, num_x, num_y, span_x, span_y, tar_y, dmin, loadx, loady,
                 truss_type="roof", support_case=1, topo_code=None):
        self._configure(num_x, num_y, span_x, span_y, tar_y, dmin, loadx, loady, truss_type, support_case, topo_code)
        print("------------------------")
        print(self.y_max)
        print(self.y_min)
        print(self.d_min)
        print("------------------------")
        self.model = None
        self.gennode()
        self.generate()

    def re_value(self, num_x, num_y, span_x, span_y, tar_y, dmin, loadx, loady,
                 truss_type="roof", support_case=1, topo_code=None):
        self._configure(num_x, num_y, span_x, span_y, tar_y, dmin, loadx, loady, truss_type, support_case, topo_code)
        self.model = None
        self.gennode()
        self.generate()

    def _configure(self, num_x, num_y, span_x, span_y, tar_y, dmin, loadx, loady, truss_type, support_case, topo_code):
        if num_y != 2:
            raise ValueError("libtfem builds two-chord trusses (num_y == 2) only")
        self.num_x, self.num_y = num_x, num_y
        self.span_x, self.span_y, self.tar_y = span_x, span_y, tar_y
        self.YoungM = 2 * 1e11
        self.truss_path = "./section_data/01_brace_rod2.csv"
        if os.path.exists(self.truss_path):
            self.truss = read_section(self.truss_path)
        else:
            self.truss = np.array(list(zip(SECTION_AREA_CM2, SECTION_INERTIA_CM4)), dtype=float)
        self.max_truss_A = self.truss[-1][0] * 1e-4
        self.max_truss_i = self.truss[-1][1] * 1e-8
        self.loadx, self.loady = loadx, loady
        self.truss_type, self.topo_code, self.support_case = truss_type, topo_code, support_case
        self.max_poss_brace_vol = 0
        self.max_short_stress = 235 * 1000000
        self.max_long_stress = 235 * 1000000 / 1.5
        self.max_deformation = 0.001 * sum(self.span_x)
        self.y_max = span_y[0]
        self.y_min = 0
        self.d_min = dmin
        sc = support_case if support_case in (1, 2, 3, 4) else 1
        self._spec = FamilySpec("dropin", num_x, tuple(span_x), tuple(span_y), tuple(tar_y), dmin, loadx, loady,
                                truss_type, sc, _backend.symmetry_of(self.SYMMETRY, num_x),
                                tuple(self.truss[:, 0]), tuple(self.truss[:, 1]))
        self._tfem = _backend.Backend(self._spec)

    # ---- mesh ----------------------------------------------------------------------------------------
    def gennode(self):
        self.n_u_x = [sum(self.span_x[:i]) for i in range(self.num_x)]
        self.n_u_y = [sum(self.span_y[:i]) for i in range(self.num_y)]
        self.n_u_coord = [[x, y] for y in self.n_u_y for x in self.n_u_x]

    def generate(self):
        t = self._tfem.tab
        nodes = []
        for i, (x, y) in enumerate(self.n_u_coord):
            n = Node()
            n.set_name(i + 1)
            n.set_coord(x, y)
            n.top_node = int(t["top"][i])
            nodes.append(n)
        k = 0
        for i, n in enumerate(nodes):
            n.vertical_pair.append(nodes[int(t["pair"][i])])
            if n.top_node == 1:
                n.target = self.tar_y[k]
                k += 1
            if t["res"][i, 0] or t["res"][i, 1]:
                n.set_res(int(t["res"][i, 0]), int(t["res"][i, 1]))
        self.model = Model()
        self.model._tfem = self._tfem
        l1 = Load()
        l1.set_name(1)
        l1.set_size(0, self.loady)
        self.model.add_load(l1)
        for n in nodes:
            self.model.add_node(n)
        last = len(self.truss) - 1
        for e, (a, b) in enumerate(t["conn"]):
            el = Element()
            el.set_name(e + 1)
            el.set_nodes(nodes[int(a)], nodes[int(b)])
            el.section_no = last
            el.set_em(self.YoungM)
            el.set_area(self.truss[last][0] * 1e-4)
            el.set_i(self.truss[last][1] * 1e-8)
            self.model.add_element(el)
        for i, n in enumerate(nodes):
            if t["loaded"][i]:
                n.set_load(l1)
                n.has_loady = 1
        self.n_u_name_div = [nodes[:self.num_x], nodes[self.num_x:]]
        self.model.restore()
        self.model.gen_all()

    # ---- move range (truss2D_GEN.py:118-133): the same expressions on the same scalar types --------------
    def set_moveRange(self):
        for n in self.model.nodes:
            pair_y = n.vertical_pair[0].coord[1]
            if n.top_node == 1:
                n.max_up = abs(self.y_max - n.coord[1])
                n.max_down = abs(n.coord[1] - pair_y - self.d_min)
            elif self.truss_type == "bridge":
                n.max_up = 0
                n.max_down = 0
            elif self.truss_type == "roof":
                n.max_up = abs(pair_y - n.coord[1] - self.d_min)
                n.max_down = abs(n.coord[1] - self.y_min)

    # ---- gene-vector objective of the MOEA/D benchmark (zip truss2D_GEN.py:117-228) -----------------------------
    def read_genes(self, genes, int_obj1, int_obj2):
        """one individual: decodes the genes onto ``self.model`` (heights, sections), analyses it and returns
        ``[obj1/int_obj1, obj2/int_obj2, con1, con2]`` like the reference; the population-sized entry is
        ``mop_truss_marl_b200.genes.GeneEvaluator.read_genes``"""
        max_height = 8 if self.num_x == 8 else (6 if self.num_x == 16 else self.y_max)
        res = self._tfem.read_genes(np.asarray(genes, dtype=np.float64), max_height, float(int_obj1), float(int_obj2))
        for n, yv in zip(self.model.nodes, res["y"][0]):
            n.coord[1] = int(yv) if float(yv) == 0.0 else float(yv)
        for el, s in zip(self.model.elements, res["section"][0]):
            el.section_no = int(s)
            el.area = self.truss[el.section_no][0] * 1e-4
            el.set_i(self.truss[el.section_no][1] * 1e-8)
        self._tfem.fill_results(self.model, res)
        return [np.float32(v) for v in res["point"][0]]

    # ---- structure text format (truss2D_GEN.py:193-211; parsed by render/truss2D_READ.py:136-172) --------
    def savetxt(self, name):
        """same lines as the reference; numbers are printed the way its pinned NumPy 1.23 prints them (float32 heights
        as ``3.2``, never ``np.float32(3.2)``), see mop_truss_marl_b200/structure_text.py"""
        from mop_truss_marl_b200.structure_text import _num
        with open(name, "w+", newline="") as f:
            for ld in self.model.loads:
                f.write(" {}, [{}, {}]\r\n".format(ld.name, _num(ld.size[0]), _num(ld.size[1])))
            for n in self.model.nodes:
                loads = "[" + ", ".join("[{}, [{}, {}]]".format(l[0].name, _num(l[0].size[0]), _num(l[0].size[1]))
                                        for l in n.loads) + "]"
                f.write(" {}, [{}, {}], [{}, {}], {}\r\n".format(n.name, _num(n.coord[0]), _num(n.coord[1]),
                                                                 _num(n.res[0]), _num(n.res[1]), loads))
            for el in self.model.elements:
                f.write(" {},{},{},{},{},[[{}]]\r\n".format(el.name, el.nodes[0].name, el.nodes[1].name, _num(el.em),
                                                            _num(el.area), _num(el.i[0][0])))

    def read_src(self, src):
        """``gen_model.read_src`` of the renderers (render/truss2D_READ.py:136-172): loads, node coordinates / supports
        and element E / A / I / section number from a structure text file onto ``self.model``"""
        from mop_truss_marl_b200.structure_text import parse_structure
        with open(src, newline="") as f:
            d = parse_structure(f.read())
        for ld in self.model.loads:
            if ld.name in d["loads"]:
                ld.size[0], ld.size[1] = d["loads"][ld.name]
        for n in self.model.nodes:
            if n.name in d["nodes"]:
                x, y, rx, ry = d["nodes"][n.name]
                n.coord[0], n.coord[1], n.res[0], n.res[1] = x, y, rx, ry
        for el in self.model.elements:
            if el.name in d["elements"]:
                _, _, em, area, inertia = d["elements"][el.name]
                el.em, el.area = em, area
                el.i[0][0] = inertia
                for k in range(len(self.truss)):
                    if el.area == self.truss[k][0] * 1e-4:
                        el.section_no = k
