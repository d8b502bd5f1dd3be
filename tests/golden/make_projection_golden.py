"""Generates tests/golden/projection_golden.npz by running the REFERENCE's own
``efficient_collaboration_weighted_projected_graph2`` (/root/reference/data.py:226-250) on seeded
@-mention graphs built the way ``DataLoader.get_graph`` builds them (data.py:302-372).

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_projection_golden.py
data.py as a whole needs pandas / sklearn / Python 2 idioms at import time, so only the one function is
taken: its source is cut out of the reference file with ``ast`` (never copied into this repository) and
executed with networkx + logging in scope.  The celebrity filter (data.py:364-370) is inline code of
get_graph and is applied here with the same three statements.
"""
import ast
import logging
import os

import networkx as nx
import numpy as np

REF = "/root/reference/data.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "projection_golden.npz")
FN = "efficient_collaboration_weighted_projected_graph2"


def load_reference_function():
    src = open(REF).read()
    tree = ast.parse(src)
    node = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == FN)
    code = compile(ast.Module(body=[node], type_ignores=[]), REF, "exec")
    ns = {"nx": nx, "logging": logging}
    exec(code, ns)
    return ns[FN]


def mention_graph(rng, n_targets, n_names, mentions_per_user, p_target):
    """data.py:305-362: target ids 0..N-1 with self loops, mentioned names get ids >= N; an edge joins a
    user and every name (or user) it mentions."""
    g = nx.Graph()
    g.add_nodes_from(range(n_targets))
    for t in range(n_targets):
        g.add_edge(t, t)                                  # data.py:309-310
    pop = 1.0 / np.arange(1, n_names + 1) ** 1.1
    pop /= pop.sum()
    for u in range(n_targets):
        k = rng.poisson(mentions_per_user)
        for _ in range(k):
            if rng.rand() < p_target:
                m = int(rng.randint(0, n_targets))
            else:
                m = n_targets + int(rng.choice(n_names, p=pop))
            g.add_edge(m, u)                              # data.py:326-327
    return g


def cases():
    rng = np.random.RandomState(77)
    return {
        "small": (mention_graph(rng, 60, 40, 2.0, 0.3), 60, 5),
        "names_only": (mention_graph(rng, 200, 150, 3.0, 0.0), 200, 10),
        "direct_only": (mention_graph(rng, 150, 1, 2.0, 1.0), 150, 10),
        "celebrity_heavy": (mention_graph(rng, 400, 30, 4.0, 0.2), 400, 8),
        "sparse": (mention_graph(rng, 500, 2000, 0.7, 0.1), 500, 10),
    }


def main():
    fn = load_reference_function()
    out = {}
    for name, (g, n, thr) in cases().items():
        M = max(g.nodes()) + 1
        e = np.array([(a, b) for a, b in g.edges()], dtype=np.int64).reshape(-1, 2)
        out[name + "__edges"] = e
        out[name + "__meta"] = np.array([n, M, thr], dtype=np.int64)
        for tag, graph in (("raw", g.copy()), ("filtered", g.copy())):
            if tag == "filtered":                         # data.py:364-370
                celebrities = []
                for i in range(n, M):
                    if i not in graph:
                        continue
                    deg = len(graph[i])
                    if deg == 1 or deg > thr:
                        celebrities.append(i)
                graph.remove_nodes_from(celebrities)
            proj = fn(graph, range(n))                    # data.py:372
            pe = np.array(sorted((min(a, b), max(a, b)) for a, b in proj.edges()), dtype=np.int64).reshape(-1, 2)
            assert set(proj.nodes()) == set(range(n))
            out["%s__proj_%s" % (name, tag)] = pe
            print(name, tag, "nodes", M, "mention edges", len(e), "projected edges", len(pe))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
