"""Host-side decoding of the CTA-per-tree node pools into the canonical DFS serialisation
(kind, depth, count, total, terminal, fanout, action) used by the parity tests and the node proxies."""
from __future__ import annotations

import numpy as np

_NAMES = ["s_total", "s_term_reward", "s_ns", "s_flags", "s_coff", "s_ccap", "s_ccnt", "s_state", "a_total", "a_value",
          "a_na", "a_visits", "a_child", "ipool", "p_s", "p_a", "p_r", "p_flag", "meta", "p_g", "a_value2", "s_next", "end"]
_DT = {"s_total": "<f8", "s_term_reward": "<f8", "s_ns": "<i4", "s_flags": "<i4", "s_coff": "<i4", "s_ccap": "<i4",
       "s_ccnt": "<i4", "a_total": "<f8", "a_value": "<f8", "a_na": "<i4", "a_visits": "<i4", "a_child": "<i4",
       "ipool": "<i4", "p_s": "<i4", "p_a": "<i4", "p_r": "<f8", "p_flag": "<i4", "meta": "<i4", "p_g": "<f8", "a_value2": "<f8",
       "s_next": "<i4"}


def read_pool(engine, tree: int) -> dict:
    L = engine.layout
    off = engine.pool_offsets()
    raw = engine._ws_bytes(L.pool_offset + tree * L.pool_stride, L.pool_stride)
    out = {}
    for i, name in enumerate(_NAMES[:-1]):
        chunk = raw[off[i]: off[i + 1]]
        if name == "s_state":
            dt = np.float32 if L.real_bytes == 4 else np.float64
            n = (len(chunk) // (dt().itemsize * L.state_dim)) * L.state_dim
            out[name] = chunk[: n * dt().itemsize].view(dt).reshape(-1, L.state_dim)
        else:
            dt = np.dtype(_DT[name])
            out[name] = chunk[: (len(chunk) // dt.itemsize) * dt.itemsize].view(dt)
    return out


def export_pool(engine, tree: int = 0) -> dict:
    P = read_pool(engine, tree)
    kinds, depths, counts, totals, terms, fans, acts, acts2, sidx = [], [], [], [], [], [], [], [], []
    two = engine.action_dim == 2
    stack = [(0, 0, 0)]  # (kind, index, depth)
    while stack:
        kind, idx, d = stack.pop()
        if kind == 0:
            n = int(P["s_ccnt"][idx]); off = int(P["s_coff"][idx])
            kinds.append(0); depths.append(d); counts.append(int(P["s_ns"][idx])); totals.append(float(P["s_total"][idx]))
            terms.append(int(P["s_flags"][idx] & 1)); fans.append(n); acts.append(0.0); acts2.append(0.0); sidx.append(idx)
            for i in range(n - 1, -1, -1):
                stack.append((1, int(P["ipool"][off + i]), d))
        else:
            # the action node's children dict: a_child = first entry, s_next = next (insertion order)
            kids = []
            c = int(P["a_child"][idx])
            while c >= 0:
                kids.append(c)
                c = int(P["s_next"][c])
            kinds.append(1); depths.append(d); counts.append(int(P["a_na"][idx])); totals.append(float(P["a_total"][idx]))
            terms.append(0); fans.append(len(kids)); acts.append(float(P["a_value"][idx])); sidx.append(-1)
            acts2.append(float(P["a_value2"][idx]) if two else 0.0)
            for c in reversed(kids):
                stack.append((0, c, d + 1))
    si = np.asarray(sidx, np.int64)
    st = np.zeros((len(si), P["s_state"].shape[1]), dtype=np.float64)
    st[si >= 0] = P["s_state"][si[si >= 0]]          # env state of every state record (action records: 0)
    return dict(kind=np.asarray(kinds, np.uint8), depth=np.asarray(depths, np.int32), count=np.asarray(counts, np.int64),
                total=np.asarray(totals, np.float64), terminal=np.asarray(terms, np.uint8),
                fanout=np.asarray(fans, np.int32), action=np.asarray(acts, np.float64), action2=np.asarray(acts2, np.float64),
                state=st)
