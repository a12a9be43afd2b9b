#!/usr/bin/env python3
"""Commonroad-free scenario front-end: CommonRoad 2018b XML -> the optimizer's input schema.

Run HERE (the build container) only; it reads /root/reference/scenarios/*.xml and
/root/reference/test/config_files/*.yaml and writes the small JSON that ships with the
package (`<pkg>/data/scenarios.json`).  Nothing at run time on the GPU box reads /root/reference.

It restates what `Configuration.find_reference_path_and_desired_velocity` does
(/root/reference/MPC_Planner/configuration.py:499-552) without commonroad:

  route reference path  -> `route_reference_path`: centre lines ((left+right)/2 per vertex) of the route lanelets, resampled,
                           lane changes ramped, smoothed -- the route planner's construction, pinned by the recorded deviation.txt
  clip_reference_path   -> configuration.py:584-623 (restated in `clip_reference_path`)
  desired_velocity      -> configuration.py:538-544 (length / ((T_end-1)*dt), rounded up to 1e-4)
  chaikins_corner_cutting + resample_polyline(step=v_des*dt) -> configuration.py:547-549
      (semantics of commonroad_dc.geometry.util recalled; package not available)
  compute_orientation_from_polyline -> configuration.py:447

Scenarios without a reference config (USA_Peach, ZAM_Tutorial-1_2, ZAM_Tutorial_Urban-3_2) are
synthesised as SURVEY.md §8(d) config 5 prescribes: path = centre line of the lanelet chain starting at
the lanelet containing x0, LF-ZAM weights, no obstacle, v_des from the clipped length (or v0 when there is no goal).
"""
import json
import math
import os
import sys
import xml.etree.ElementTree as ET

import numpy as np
import yaml

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..",
                   "motion-planning-for-autonomous-driving-with-mpc_b200", "data", "scenarios.json")


def _pts(node):
    return np.array([[float(p.find("x").text), float(p.find("y").text)] for p in node.findall("point")])


def _exact(node, default=None):
    if node is None:
        return default
    e = node.find("exact")
    if e is not None:
        return float(e.text)
    return default


def parse_xml(path):
    root = ET.parse(path).getroot()
    dt = float(root.attrib.get("timeStepSize", 0.1))
    lanelets = {}
    for ln in root.findall("lanelet"):
        lid = int(ln.attrib["id"])
        left, right = _pts(ln.find("leftBound")), _pts(ln.find("rightBound"))
        lanelets[lid] = dict(left=left, right=right, center=0.5 * (left + right),
                             succ=[int(s.attrib["ref"]) for s in ln.findall("successor")],
                             pred=[int(s.attrib["ref"]) for s in ln.findall("predecessor")],
                             adj=[int(a.attrib["ref"]) for tag in ("adjacentLeft", "adjacentRight")
                                  for a in ln.findall(tag) if a.attrib.get("drivingDir") == "same"])
    obstacles = []
    for ob in root.findall("obstacle"):
        role = ob.find("role").text
        shape = ob.find("shape")
        rect = shape.find("rectangle") if shape is not None else None
        st = ob.find("initialState")
        if rect is None or st is None:
            continue
        pos = st.find("position").find("point")
        obstacles.append(dict(role=role, length=float(rect.find("length").text), width=float(rect.find("width").text),
                              x=float(pos.find("x").text), y=float(pos.find("y").text),
                              orientation=_exact(st.find("orientation"), 0.0)))
    pp = root.find("planningProblem")
    ist = pp.find("initialState")
    pos = ist.find("position").find("point")
    init = dict(x=float(pos.find("x").text), y=float(pos.find("y").text), v=_exact(ist.find("velocity")),
                psi=_exact(ist.find("orientation")))
    goal = dict(center=None, lanelets=[], t_end=None)
    gs = pp.find("goalState")
    if gs is not None:
        gp = gs.find("position")
        if gp is not None:
            rect = gp.find("rectangle")
            if rect is not None:
                c = rect.find("center")
                goal["center"] = [float(c.find("x").text), float(c.find("y").text)]
            goal["lanelets"] = [int(l.attrib["ref"]) for l in gp.findall("lanelet")]
        tt = gs.find("time")
        if tt is not None:
            if tt.find("intervalEnd") is not None:
                goal["t_end"] = int(float(tt.find("intervalEnd").text))
            else:
                goal["t_end"] = int(_exact(tt))
    return dict(dt=dt, lanelets=lanelets, obstacles=obstacles, init=init, goal=goal, pp_id=int(pp.attrib["id"]))


# ---------------------------------------------------------------- geometry (commonroad_dc.geometry.util restated)
def chaikins_corner_cutting(P, refinements=1):
    P = np.asarray(P, float)
    for _ in range(refinements):
        out = [P[0]]
        for a, b in zip(P[:-1], P[1:]):
            out.append(0.75 * a + 0.25 * b)
            out.append(0.25 * a + 0.75 * b)
        out.append(P[-1])
        P = np.array(out)
    return P


def resample_polyline(P, step):
    P = np.asarray(P, float)
    out = [P[0]]
    current_position = step
    current_length = np.linalg.norm(P[0] - P[1])
    current_idx = 0
    while current_idx < len(P) - 1:
        if current_position >= current_length:
            current_position -= current_length
            current_idx += 1
            if current_idx > len(P) - 2:
                break
            current_length = np.linalg.norm(P[current_idx + 1] - P[current_idx])
        else:
            rel = current_position / current_length
            out.append((1 - rel) * P[current_idx] + rel * P[current_idx + 1])
            current_position += step
    if np.linalg.norm(out[-1] - P[-1]) >= 1e-6:
        out.append(P[-1])
    return np.array(out)


def compute_orientation_from_polyline(P):
    d = np.diff(P, axis=0)
    o = np.arctan2(d[:, 1], d[:, 0])
    return np.concatenate([o, o[-1:]])


def polyline_length(P):
    return float(np.sum(np.linalg.norm(np.diff(P, axis=0), axis=1)))


def find_closest_point(path, p):  # configuration.py:26-37
    d = path - p.reshape(1, 2)
    return int(np.argmin((d ** 2).sum(1)))


def clip_reference_path(path, init_position, goal_position):  # configuration.py:584-623
    si, ei = find_closest_point(path, init_position), find_closest_point(path, goal_position)
    if goal_position[0] >= init_position[0]:
        if ((path[si] - init_position) >= 0).sum() != 2:
            si += 1
        if ((path[ei] - goal_position) <= 0).sum() != 2:
            ei -= 1
    else:
        if ((path[si] - init_position) <= 0).sum() != 2:
            si += 1
        if ((path[ei] - goal_position) >= 0).sum() != 2:
            ei -= 1
    return np.concatenate([init_position.reshape(1, 2), path[si:ei + 1], goal_position.reshape(1, 2)], axis=0)


# ---------------------------------------------------------------- route (lanelet graph search, centre lines)
def _point_in_lanelet(ln, p):
    poly = np.concatenate([ln["left"], ln["right"][::-1]])
    x, y = p
    inside = False
    n = len(poly)
    for i in range(n):
        x1, y1 = poly[i]
        x2, y2 = poly[(i + 1) % n]
        if (y1 > y) != (y2 > y):
            if x < (x2 - x1) * (y - y1) / (y2 - y1 + 1e-300) + x1:
                inside = not inside
    return inside


def start_lanelets(sc):
    p = np.array([sc["init"]["x"], sc["init"]["y"]])
    psi = sc["init"]["psi"]
    cands = []
    for lid, ln in sc["lanelets"].items():
        if _point_in_lanelet(ln, p):
            c = ln["center"]
            i = find_closest_point(c, p)
            j = min(i + 1, len(c) - 1)
            i0 = j - 1
            h = math.atan2(c[j, 1] - c[i0, 1], c[j, 0] - c[i0, 0])
            dpsi = abs((h - psi + math.pi) % (2 * math.pi) - math.pi)
            cands.append((dpsi, lid))
    if not cands:  # nearest centre line
        best = min(sc["lanelets"].items(), key=lambda kv: np.min(((kv[1]["center"] - p) ** 2).sum(1)))
        return [best[0]]
    return [lid for _, lid in sorted(cands)]


def route_lanelets(sc):
    goals = set(sc["goal"]["lanelets"])
    if sc["goal"]["center"] is not None:
        g = np.array(sc["goal"]["center"])
        goals |= {lid for lid, ln in sc["lanelets"].items() if _point_in_lanelet(ln, g)}
    for s in start_lanelets(sc):
        # BFS over successors
        prev = {s: None}
        queue = [s]
        found = None
        while queue:
            cur = queue.pop(0)
            if cur in goals:
                found = cur
                break
            for nx in sc["lanelets"][cur]["succ"] + sc["lanelets"][cur]["adj"]:
                if nx in sc["lanelets"] and nx not in prev:
                    prev[nx] = cur
                    queue.append(nx)
        if found is not None or not goals:
            if found is None:  # no goal: follow first successors
                chain = [s]
                while sc["lanelets"][chain[-1]]["succ"] and len(chain) < 8:
                    nx = sc["lanelets"][chain[-1]]["succ"][0]
                    if nx in chain or nx not in sc["lanelets"]:
                        break
                    chain.append(nx)
                return chain
            chain = []
            while found is not None:
                chain.append(found)
                found = prev[found]
            return chain[::-1]
    return [start_lanelets(sc)[0]]


def route_reference_path(sc, chain, step_resample=1.0, num_vertices_lane_change_max=6, percentage_vertices_lane_change_max=0.1,
                         step_final=2.0, refinements=4):
    """The route planner's reference path for a lanelet route, restated (commonroad-route-planner is not installed here):

      1. every lanelet's centre line is resampled at 1 m;
      2. a run of lane changes (next lanelet laterally adjacent instead of a successor, e.g. Lanker 3452 -> 3454 -> 3456) shares
         the run's length evenly: lanelet k of n contributes the vertices of its k-th n-th, minus a few vertices
         (min(int(0.1 * n_vertices) + 1, 6)) at each junction so that the change is a ramp, not a step;
      3. the concatenation is resampled at 2 m and smoothed by four rounds of Chaikin corner cutting.

    PINNED BY THE REFERENCE'S OWN RECORDINGS: `deviation.txt` (mpc_planner.py:190-197) is the distance of every recorded state to
    the closest VERTEX of exactly this path.  With the constants above all six recorded files (ZAM_Over lane following /
    collision avoidance and USA_Lanker lane following, CasADi and Forcespro runs, 260 values) are reproduced to < 1e-4
    (tests/test_results_format.py); any other resampling step, vertex allowance or number of Chaikin rounds misses by 0.03 m or
    more (scan in profiles/r02_summary.md)."""
    L = sc["lanelets"]
    instr = [1 if (b in L[a]["adj"] and b not in L[a]["succ"]) else 0 for a, b in zip(chain[:-1], chain[1:])] + [0]
    portions = [None] * len(chain)
    i = 0
    while i < len(chain):
        if instr[i] == 0:
            portions[i] = (0.0, 1.0)
            i += 1
        else:
            j = i
            while instr[j] == 1:
                j += 1
            n = j - i + 1                                  # lanelets i..j share [0, 1]
            for k in range(n):
                portions[i + k] = (k / n, (k + 1) / n)
            i = j + 1
    ref = None
    for idx, lid in enumerate(chain):
        v = resample_polyline(L[lid]["center"], step_resample)
        nv = len(v)
        nlc = min(int(nv * percentage_vertices_lane_change_max) + 1, num_vertices_lane_change_max)
        last = idx == len(chain) - 1
        if ref is None:
            i0, i1 = int(portions[idx][0] * nv), max(int(portions[idx][1] * nv), 1)
            if not last:
                i1 = max(i1 - nlc, 1)
            ref = v[i0:i1]
        else:
            i0, i1 = min(int(portions[idx][0] * nv) + nlc, nv - 1), int(portions[idx][1] * nv)
            if not last:
                i1 = max(i1 - nlc, 1)
            ref = np.concatenate([ref, v[i0:i1]])
    return chaikins_corner_cutting(resample_polyline(ref, step_final), refinements)


# ---------------------------------------------------------------- weights
LF_ZAM_YAML = "test/config_files/config_LF_ZAM_Over-1_1.yaml"
CA_ZAM_YAML = "test/config_files/config_CA_ZAM_Over-1_1.yaml"
LF_LANKER_YAML = "test/config_files/config_LF_USA_Lanker-2_18_T-1.yaml"


def load_yaml(rel):
    with open(os.path.join(REF, rel)) as f:
        return yaml.safe_load(f)


def build(name, xml, yaml_rel, use_case, synth=False):
    sc = parse_xml(os.path.join(REF, "scenarios", xml))
    settings = load_yaml(yaml_rel)
    chain = route_lanelets(sc)
    origin = route_reference_path(sc, chain)
    init_pos = np.array([sc["init"]["x"], sc["init"]["y"]])
    goal_pos = np.array(sc["goal"]["center"]) if sc["goal"]["center"] is not None else origin[-1].copy()
    clipped = clip_reference_path(origin, init_pos, goal_pos)
    dt = sc["dt"]
    length = polyline_length(clipped)
    t_end = sc["goal"]["t_end"]
    if t_end is None:  # synthesised (no goal state): keep the initial speed, T = N+1 = 31 points
        v_des = sc["init"]["v"]
        t_end = 31
        keep = v_des * dt * (t_end - 1)
        # truncate the clipped path to the needed length
        acc = np.concatenate([[0], np.cumsum(np.linalg.norm(np.diff(clipped, axis=0), axis=1))])
        n_keep = int(np.searchsorted(acc, keep + 1e-9)) + 1
        clipped = clipped[:max(n_keep, 2)]
        length = polyline_length(clipped)
    v_des = length / ((t_end - 1) * dt)
    v_des = round(v_des, 4) + 0.0001 if v_des > round(v_des, 4) else round(v_des, 4)
    path = resample_polyline(chaikins_corner_cutting(clipped), step=v_des * dt)
    orient = compute_orientation_from_polyline(path)
    if use_case == "collision_avoidance":
        ob = sc["obstacles"][0]
        obstacle = dict(position_x=ob["x"], position_y=ob["y"], length=ob["length"], width=ob["width"],
                        orientation=ob["orientation"])
    else:  # configuration.py:477-483
        obstacle = dict(position_x=-100.0, position_y=0.0, length=0.0, width=0.0, orientation=0.0)
    # road boundaries as Configuration sets them (configuration.py:432-433): the RIGHT vertices of the 2nd and of the 1st lanelet
    # of the network (hard-wired indices in the reference: meaningful for the two-lanelet ZAM_Over road only)
    lids = list(sc["lanelets"].keys())
    bounds = {}
    if len(lids) == 2:
        bounds = dict(left_road_boundary=sc["lanelets"][lids[1]]["right"].tolist(), right_road_boundary=sc["lanelets"][lids[0]]["right"].tolist())
    return dict(**bounds, name=name, xml=xml, use_case=use_case, synthesised=synth, dt=dt,
                x0=[sc["init"]["x"], sc["init"]["y"], 0.0, sc["init"]["v"], sc["init"]["psi"]],
                route_lanelets=chain, desired_velocity=v_des, iter_length=int(path.shape[0]),
                origin_reference_path=origin.tolist(),
                reference_path=path.tolist(), orientation=orient.tolist(),
                clipped_length=length, static_obstacle=obstacle,
                weights_setting={k: float(v) for k, v in settings["weights_setting"].items()},
                wheelbase=2.578)


def main():
    out = {}
    out["ZAM_Over-1_1_LF"] = build("ZAM_Over-1_1_LF", "ZAM_Over-1_1.xml", LF_ZAM_YAML, "lane_following")
    out["ZAM_Over-1_1_CA"] = build("ZAM_Over-1_1_CA", "ZAM_Over-1_1.xml", CA_ZAM_YAML, "collision_avoidance")
    out["USA_Lanker-2_18_T-1_LF"] = build("USA_Lanker-2_18_T-1_LF", "USA_Lanker-2_18_T-1.xml", LF_LANKER_YAML,
                                          "lane_following")
    out["ZAM_Over-1_1_LFfile"] = build("ZAM_Over-1_1_LFfile", "ZAM_Over-1_1_LF.xml", LF_ZAM_YAML, "lane_following")
    for nm, xml in (("USA_Peach-2_1_T-1", "USA_Peach-2_1_T-1.xml"),
                    ("ZAM_Tutorial-1_2_T-1", "ZAM_Tutorial-1_2_T-1.xml"),
                    ("ZAM_Tutorial_Urban-3_2", "ZAM_Tutorial_Urban-3_2.xml")):
        try:
            out[nm] = build(nm, xml, LF_ZAM_YAML, "lane_following", synth=True)
        except Exception as e:  # noqa
            print("synth failed for", nm, e, file=sys.stderr)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as f:
        json.dump(out, f)
    for k, v in out.items():
        p = np.array(v["reference_path"])
        print(f"{k:28s} T={v['iter_length']:3d} v_des={v['desired_velocity']:.4f} dt={v['dt']} route={v['route_lanelets']} "
              f"p0={p[0].round(3)} pT={p[-1].round(3)} len={v['clipped_length']:.3f}")


if __name__ == "__main__":
    main()
