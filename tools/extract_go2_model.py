#!/usr/bin/env python3
"""Extract the GO2 model constants the hot path needs from the reference MJCF.

Run HERE (container with /root/reference mounted), never on the GPU box:

    python tools/extract_go2_model.py [/root/reference] \
        > phase_guided_terrain_traversal_b200/assets/go2_model.json

Reads (reference, read-only):
    go2/xmls/go2_mjx_feetonly.xml      robot: bodies, inertials, joints, feet geoms, sites,
                                       actuators, sensors, keyframe, options, custom numerics
    go2/xmls/scene_mjx_feetonly.xml    task "flat_terrain": floor plane
    go2/xmls/terrain_scene_mjx.xml     task "stairs": floor plane + 100 placeholder boxes

Writes a JSON document in this repo's own schema (see model.py: `load_spec`). Only a small
MJCF subset is understood (nested <default> classes, childclass, <include>, radian angles,
autolimits) - exactly what those three files use. No reference text is copied; the output
is a table of numbers with the XML line each group came from.
"""
from __future__ import annotations

import json
import sys
import xml.etree.ElementTree as ET
from pathlib import Path


def _parse_with_lines(path: Path):
    """Parse XML into ElementTree elements, recording each element's source line as the
    pseudo-attribute `__line__` (used only for citations in the output)."""
    import xml.parsers.expat as expat

    parser = expat.ParserCreate()
    stack, root = [], [None]

    def start(tag, attrs):
        el = ET.Element(tag, attrs)
        el.set("__line__", str(parser.CurrentLineNumber))
        if stack:
            stack[-1].append(el)
        else:
            root[0] = el
        stack.append(el)

    def end(tag):
        stack.pop()

    parser.StartElementHandler = start
    parser.EndElementHandler = end
    with open(path, "rb") as f:
        parser.ParseFile(f)
    return root[0]


def floats(s):
    return [float(x) for x in s.split()]


class Defaults:
    """Nested MJCF default classes: class name -> {tag -> attrs}, with parent links."""

    def __init__(self):
        self.attrs = {"main": {}}
        self.parent = {"main": None}

    def load(self, node, cls="main"):
        for child in node:
            if child.tag == "default":
                name = child.get("class")
                if name is None:  # top-level <default> without class == main
                    self.load(child, cls)
                    continue
                self.attrs.setdefault(name, {})
                self.parent[name] = cls
                self.load(child, name)
            else:
                d = self.attrs[cls].setdefault(child.tag, {})
                for k, v in child.attrib.items():
                    if k != "__line__":
                        d[k] = v

    def resolve(self, tag, cls, own):
        chain = []
        c = cls or "main"
        while c is not None:
            chain.append(c)
            c = self.parent[c]
        out = {}
        for c in reversed(chain):
            out.update(self.attrs[c].get(tag, {}))
        out.update({k: v for k, v in own.items() if k not in ("class", "__line__")})
        return out


def main():
    ref = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    xml_dir = ref / "go2" / "xmls"
    robot = _parse_with_lines(xml_dir / "go2_mjx_feetonly.xml")

    defaults = Defaults()
    for d in robot.findall("default"):
        defaults.load(d)

    # ---- options (later <option> elements override earlier ones) -------------------------
    opt = {"timestep": 0.002, "impratio": 1.0, "iterations": 100, "ls_iterations": 50,
           "tolerance": 1e-8, "ls_tolerance": 0.01, "cone": "pyramidal", "integrator": "Euler",
           "gravity": [0.0, 0.0, -9.81], "eulerdamp": True}
    opt_lines = []
    for o in robot.findall("option"):
        opt_lines.append(int(o.get("__line__")))
        for k in ("timestep", "impratio", "tolerance", "ls_tolerance"):
            if o.get(k) is not None:
                opt[k] = float(o.get(k))
        for k in ("iterations", "ls_iterations"):
            if o.get(k) is not None:
                opt[k] = int(o.get(k))
        for k in ("cone", "integrator"):
            if o.get(k) is not None:
                opt[k] = o.get(k)
        for fl in o.findall("flag"):
            if fl.get("eulerdamp") == "disable":
                opt["eulerdamp"] = False
    numerics = {n.get("name"): float(n.get("data")) for n in robot.find("custom").findall("numeric")}

    comp = robot.find("compiler")
    assert comp.get("angle") == "radian" and comp.get("autolimits") == "true"

    # ---- walk the kinematic tree --------------------------------------------------------
    bodies = [{"name": "world", "parent": -1, "pos": [0, 0, 0], "quat": [1, 0, 0, 0],
               "ipos": [0, 0, 0], "iquat": [1, 0, 0, 0], "mass": 0.0, "inertia": [0, 0, 0]}]
    joints, geoms, sites = [], [], []
    ngeom = [0]

    def add_geom(g, cls, body_id):
        a = defaults.resolve("geom", g.get("class") or cls, g.attrib)
        gid = ngeom[0]
        ngeom[0] += 1
        contype = int(a.get("contype", 1))
        conaff = int(a.get("conaffinity", 1))
        if contype == 0 and conaff == 0:
            return  # visual / non-colliding geoms carry no dynamics (explicit <inertial>)
        geoms.append({
            "id": gid, "name": g.get("name"), "body": body_id, "type": a.get("type", "sphere"),
            "size": floats(a.get("size", "0")), "pos": floats(a.get("pos", "0 0 0")),
            "friction": (floats(a.get("friction", "1 0.005 0.0001")) + [0.005, 0.0001])[:3]
            if len(floats(a.get("friction", "1 0.005 0.0001"))) < 3 else floats(a.get("friction")),
            "margin": float(a.get("margin", 0)), "gap": float(a.get("gap", 0)),
            "condim": int(a.get("condim", 3)), "contype": contype, "conaffinity": conaff,
            "solref": floats(a.get("solref", "0.02 1")),
            "solimp": (floats(a.get("solimp", "0.9 0.95 0.001 0.5 2")) + [0.5, 2.0])[:5]
            if len(floats(a.get("solimp", "0.9 0.95 0.001 0.5 2"))) == 3
            else floats(a.get("solimp", "0.9 0.95 0.001 0.5 2")),
            "group": int(a.get("group", 0)), "line": int(g.get("__line__")),
        })

    def walk(node, parent_id, cls):
        for child in node:
            if child.tag == "body":
                ccls = child.get("childclass") or cls
                bid = len(bodies)
                inert = child.find("inertial")
                b = {"name": child.get("name"), "parent": parent_id,
                     "pos": floats(child.get("pos", "0 0 0")),
                     "quat": floats(child.get("quat", "1 0 0 0")),
                     "ipos": floats(inert.get("pos")), "iquat": floats(inert.get("quat", "1 0 0 0")),
                     "mass": float(inert.get("mass")), "inertia": floats(inert.get("diaginertia")),
                     "line": int(child.get("__line__"))}
                bodies.append(b)
                for j in child:
                    if j.tag == "freejoint":
                        joints.append({"name": "root", "type": "free", "body": bid,
                                       "line": int(j.get("__line__"))})
                    elif j.tag == "joint":
                        a = defaults.resolve("joint", j.get("class") or ccls, j.attrib)
                        joints.append({"name": a.get("name"), "type": a.get("type", "hinge"),
                                       "body": bid, "axis": floats(a.get("axis", "0 0 1")),
                                       "pos": floats(a.get("pos", "0 0 0")),
                                       "range": floats(a["range"]),
                                       "damping": float(a.get("damping", 0)),
                                       "armature": float(a.get("armature", 0)),
                                       "frictionloss": float(a.get("frictionloss", 0)),
                                       "solref_limit": floats(a.get("solreflimit", "0.02 1")),
                                       "solimp_limit": floats(a.get("solimplimit", "0.9 0.95 0.001 0.5 2")),
                                       "margin": float(a.get("margin", 0)),
                                       "line": int(j.get("__line__"))})
                    elif j.tag == "geom":
                        add_geom(j, ccls, bid)
                    elif j.tag == "site":
                        a = defaults.resolve("site", j.get("class") or ccls, j.attrib)
                        sites.append({"name": a.get("name"), "body": bid,
                                      "pos": floats(a.get("pos", "0 0 0")),
                                      "line": int(j.get("__line__"))})
                walk(child, bid, ccls)

    # Scene files <include> the robot FIRST, then add world geoms/bodies; but MuJoCo orders
    # geoms by body, and world-body geoms (the floor) therefore get the lowest ids.
    scenes = {}
    for task, fname in (("flat_terrain", "scene_mjx_feetonly.xml"), ("stairs", "terrain_scene_mjx.xml")):
        sc = _parse_with_lines(xml_dir / fname)
        assert sc.find("include").get("file") == "go2_mjx_feetonly.xml"
        wb = sc.find("worldbody")
        floor = [g for g in wb.findall("geom") if g.get("name") == "floor"][0]
        boxes = []
        for b in wb.findall("body"):
            g = b.find("geom")
            boxes.append({"name": b.get("name"), "pos": floats(b.get("pos")), "quat": floats(b.get("quat")),
                          "size": floats(g.get("size")), "contype": int(g.get("contype")),
                          "conaffinity": int(g.get("conaffinity"))})
        scenes[task] = {
            "file": f"go2/xmls/{fname}",
            "floor": {"type": floor.get("type"), "pos": floats(floor.get("pos")),
                      "contype": int(floor.get("contype")), "conaffinity": int(floor.get("conaffinity")),
                      "friction": [1.0, 0.005, 0.0001], "solref": [0.02, 1.0],
                      "solimp": [0.9, 0.95, 0.001, 0.5, 2.0], "margin": 0.0, "condim": 3,
                      "line": int(floor.get("__line__"))},
            "n_boxes": len(boxes),
            # every placeholder is identical up to its parking position (100+k,100+k,10)
            "box_template": ({"size": boxes[0]["size"], "quat": boxes[0]["quat"],
                              "contype": boxes[0]["contype"], "conaffinity": boxes[0]["conaffinity"],
                              "pos0": boxes[0]["pos"], "pos_step": [b - a for a, b in zip(boxes[0]["pos"], boxes[1]["pos"])],
                              "friction": [1.0, 0.005, 0.0001], "solref": [0.02, 1.0],
                              "solimp": [0.9, 0.95, 0.001, 0.5, 2.0], "margin": 0.0, "condim": 3}
                             if boxes else None),
        }
        if boxes:
            for k, b in enumerate(boxes):
                assert b["size"] == boxes[0]["size"] and b["quat"] == boxes[0]["quat"]
                assert b["pos"] == [boxes[0]["pos"][0] + k, boxes[0]["pos"][1] + k, boxes[0]["pos"][2]]

    # world-body geoms come first in MuJoCo's geom numbering: floor = geom 0
    ngeom[0] = 1
    walk(robot.find("worldbody"), 0, None)
    n_robot_geoms_end = ngeom[0]

    # ---- actuators ----------------------------------------------------------------------
    actuators = []
    for a_el in robot.find("actuator"):
        assert a_el.tag == "position"
        a = defaults.resolve("general", a_el.get("class"), a_el.attrib)
        gain = floats(a.get("gainprm", "1 0 0"))
        bias = floats(a.get("biasprm", "0 0 0"))
        # <position>: kp (if given) -> gainprm[0]; biasprm[1] = -gainprm[0]; kv (if given) -> biasprm[2] = -kv;
        # an absent kv leaves biasprm[2] at the class default.
        if "kp" in a:
            gain[0] = float(a["kp"])
        bias[1] = -gain[0]
        if "kv" in a:
            bias[2] = -float(a["kv"])
        actuators.append({"name": a.get("name"), "joint": a.get("joint"),
                          "gainprm": gain[:3], "biasprm": bias[:3],
                          "ctrlrange": floats(a["ctrlrange"]), "forcerange": floats(a["forcerange"]),
                          "gear": 1.0, "line": int(a_el.get("__line__"))})

    sensors = []
    adr = 0
    dims = {"gyro": 3, "accelerometer": 3, "framequat": 4, "framepos": 3, "framelinvel": 3,
            "frameangvel": 3, "velocimeter": 3, "framezaxis": 3}
    for s in robot.find("sensor"):
        sensors.append({"name": s.get("name"), "type": s.tag, "adr": adr, "dim": dims[s.tag],
                        "obj": s.get("site") or s.get("objname"), "ref": s.get("refname"),
                        "line": int(s.get("__line__"))})
        adr += dims[s.tag]

    key = robot.find("keyframe").find("key")
    spec = {
        "schema": "pgtt-b200/go2-model/1",
        "source": {"robot": "go2/xmls/go2_mjx_feetonly.xml", "option_lines": opt_lines,
                   "note": "numbers transcribed by tools/extract_go2_model.py; base.py:57-62 overrides "
                           "(timestep, Kp, Kd) are applied at run time from the config, not here"},
        "option": opt, "numeric": numerics,
        "bodies": bodies, "joints": joints, "geoms": geoms, "sites": sites,
        "actuators": actuators, "sensors": sensors, "nsensordata": adr,
        "keyframe_home": {"qpos": floats(key.get("qpos")), "ctrl": floats(key.get("ctrl")),
                          "line": int(key.get("__line__"))},
        "n_robot_geoms_end": n_robot_geoms_end,
        "scenes": scenes,
    }
    json.dump(spec, sys.stdout, indent=1)
    sys.stdout.write("\n")


if __name__ == "__main__":
    main()
