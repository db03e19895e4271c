"""Drop-in for the reference's ``FEM_2Dtruss`` module: the same container classes and attribute names
(``Load``, ``Node``, ``Element``, ``Model``; reference ``FEM_2Dtruss.py:12-161``), with ``Model.gen_all`` solved
on the GPU through ``tfem_solve_only`` instead of the NumPy direct-stiffness loops (``:434-459``).

Only models built by the drop-in ``truss2D_GEN.gen_model`` (the two-chord truss families libtfem knows) can be
solved; anything else raises -- there is no CPU solver behind this module."""
import numpy as np


class Load:
    def __init__(self):
        self.name = 1
        self.size = [0, 0]

    def set_name(self, name):
        self.name = name

    def set_size(self, x, y):
        self.size[0], self.size[1] = x, y

    def __repr__(self):
        return "{0}, {1}".format(self.name, self.size)


class Node:
    def __init__(self):
        self.name = 1
        self.coord = [0, 0]
        self.res = [0, 0]
        self.loads = []
        self.global_d = []
        self.adj_ele = []
        self.connected = 0
        self.top_node = 0
        self.vertical_pair = []
        self.int_y = 0
        self.max_up = 0
        self.max_down = 0
        self.target = 0
        self.has_loady = 0

    def set_target(self):
        if self.top_node == 0:
            self.target = self.coord[1]

    def set_name(self, name):
        self.name = name

    def set_coord(self, xval, yval):
        self.coord[0], self.coord[1] = xval, yval
        self.int_y = yval

    def set_res(self, xres, yres):
        self.res[0], self.res[1] = xres, yres

    def set_load(self, load):
        self.loads.append([load])
        self.has_loady = load.size[1]

    def __repr__(self):
        return "{0}, {1}, {2}, {3}".format(self.name, self.coord, self.res, self.loads)


class Element(Node):
    def __init__(self):
        self.name = 1
        self.nodes = []
        self.em = 0
        self.area = 0
        self.dia = 0
        self.length = None
        self.e_q = []
        self.i = [[0]]
        self.section_no = 0
        self.has_changed = 0
        self.yield_stress = 235 * 1e6
        self.long_stress = self.yield_stress / 1.5
        self.iscompress = None
        self.prop_yeield = 0

    def gen_length(self):
        dx = self.nodes[1].coord[0] - self.nodes[0].coord[0]
        dy = self.nodes[1].coord[1] - self.nodes[0].coord[1]
        self.length = (dx ** 2 + dy ** 2) ** 0.5
        return self.length

    def set_nodes(self, startnode, endnode):
        self.nodes += [startnode, endnode]
        for n in (startnode, endnode):
            n.adj_ele.append(self.name)
            n.connected += 1

    def set_em(self, emval):
        self.em = emval

    def set_area(self, area):
        self.area = area

    def set_i(self, xval):
        self.i[0][0] = xval

    def __repr__(self):
        return "{0}, {1}, {2}".format(self.nodes, self.em, self.area)


class Model:
    _RESULT_FIELDS = ("jp", "pj", "nsc", "tnsc", "ttnsc", "jlv", "local_k", "global_k", "T_matrix", "Tt_matrix",
                      "ssm", "d", "v", "u", "q", "f", "r")

    def __init__(self):
        self.nodes, self.elements, self.loads = [], [], []
        self._tfem = None               # set by the drop-in gen_model: the libtfem backend of this family
        self.restore()

    def restore(self):
        for name in self._RESULT_FIELDS:
            setattr(self, name, [])
        self.ndof = 0
        self.U_full = 0

    def add_load(self, load):
        self.loads.append(load)

    def add_node(self, node):
        self.nodes.append(node)

    def add_element(self, element):
        self.elements.append(element)

    def reset(self):
        self.nodes, self.elements = [], []

    def gen_all(self):
        """solve the current geometry / sections on the GPU and fill the reference's result fields"""
        if self._tfem is None:
            raise NotImplementedError("this Model was not built by the drop-in gen_model: libtfem only solves "
                                      "the two-chord truss families and there is no CPU solver here")
        self._tfem.solve_into(self)
