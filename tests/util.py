import numpy as np

import oracle
from scintirete_b200 import GraphState


def gaussian(n, d, seed):
    return np.random.default_rng(seed).standard_normal((n, d)).astype(np.float32)


def to_graph_state(st: "oracle.GraphState", m: int) -> GraphState:
    return GraphState(st.ids, st.list_counts, st.edge_counts, st.edges, st.entrypoint, st.max_layer, st.size,
                      st.deleted, st.vectors, m=m)


def recall(ids, gt):
    k = gt.shape[1]
    return float(np.mean([len(set(ids[i]) & set(gt[i])) / k for i in range(len(gt))]))
