"""MJCF -> model digest for the fake mujoco_py (TEST INFRASTRUCTURE, see oracle/refharness/__init__.py).

The reference loads its scene with mujoco_py.load_model_from_path(<assets/fetch/Nblock(s).xml>)
(robot_env.py:24, fetch_env.py:553).  The fake needs only what the reference's Python reads from the
model -- geom names in id order (fetch_env.py:106-117,284), opt.timestep (robot_env.py:47-48,
fetch_env.py:190) -- plus the handful of geometry constants BlockPhys pins (checked against the
oracle's constants at load time).  This module extracts that digest from the real XML; oracle/build_ref.py
stores it (not the XML) next to the compiled reference so the harness also runs where /root/reference
does not exist.
"""
import json
import os
import xml.etree.ElementTree as ET


def _walk(elem, base_dir, out, depth=0):
    for child in list(elem):
        tag = child.tag
        if tag == "include":
            sub = ET.parse(os.path.join(base_dir, child.attrib["file"])).getroot()
            # an included file's root (<mujoco> or <mujocoinclude>) is spliced in place
            _walk(sub, base_dir, out, depth)
            continue
        if tag == "option" and "timestep" in child.attrib:
            out["timestep"] = float(child.attrib["timestep"])
        elif tag == "geom":
            out["geoms"].append(dict(name=child.attrib.get("name"), size=child.attrib.get("size"),
                                     type=child.attrib.get("type"), pos=child.attrib.get("pos"),
                                     body=out["_body_stack"][-1] if out["_body_stack"] else None))
        elif tag == "joint" and "name" in child.attrib:
            out["joints"].append(dict(name=child.attrib["name"], type=child.attrib.get("type", "hinge"),
                                      range=child.attrib.get("range"), body=out["_body_stack"][-1] if out["_body_stack"] else None))
        elif tag == "site" and "name" in child.attrib:
            out["sites"].append(dict(name=child.attrib["name"], pos=child.attrib.get("pos"),
                                     body=out["_body_stack"][-1] if out["_body_stack"] else None))
        elif tag == "position":
            out["actuators"].append(dict(name=child.attrib.get("name"), joint=child.attrib.get("joint"),
                                         kp=float(child.attrib.get("kp", "1")), ctrlrange=child.attrib.get("ctrlrange")))
        elif tag == "weld":
            out["welds"].append(dict(body1=child.attrib.get("body1"), body2=child.attrib.get("body2"), solref=child.attrib.get("solref")))
        if tag == "body":
            out["bodies"].append(dict(name=child.attrib.get("name"), pos=child.attrib.get("pos"), mocap=child.attrib.get("mocap")))
            out["_body_stack"].append(child.attrib.get("name"))
            _walk(child, base_dir, out, depth + 1)
            out["_body_stack"].pop()
        else:
            _walk(child, base_dir, out, depth + 1)


def digest_from_xml(path):
    """Parse an MJCF file (following <include>) into the digest the fake MjSim consumes."""
    root = ET.parse(path).getroot()
    out = dict(timestep=0.002, geoms=[], joints=[], sites=[], bodies=[], actuators=[], welds=[], _body_stack=[])
    _walk(root, os.path.dirname(path), out)
    del out["_body_stack"]
    geom_names = [g["name"] for g in out["geoms"]]
    nblocks = sum(1 for n in geom_names if n is not None and n.startswith("object"))
    f3 = lambda s: [float(x) for x in s.split()]
    table_geom = next(g for g in out["geoms"] if g["name"] == "table")
    table_body = next(b for b in out["bodies"] if b["name"] == "table0")
    cube = next(g for g in out["geoms"] if g["name"] == "object0")
    finger = next(g for g in out["geoms"] if g["name"] == "robot0:r_gripper_finger_link")
    return dict(
        source=os.path.basename(path),
        timestep=out["timestep"],
        geom_names=geom_names,
        nblocks=nblocks,
        joint_names=[j["name"] for j in out["joints"]],
        robot_joint_names=[j["name"] for j in out["joints"] if j["name"].startswith("robot")],
        site_names=[s["name"] for s in out["sites"]],
        body_names=[b["name"] for b in out["bodies"]],
        table_half=f3(table_geom["size"]), table_body_pos=f3(table_body["pos"]),
        cube_half=f3(cube["size"]), finger_half=f3(finger["size"]),
        actuators=out["actuators"], welds=out["welds"],
    )


def load_digest(path):
    """`path` is either a real MJCF file or a digest written by oracle/build_ref.py under the same name."""
    with open(path, "rb") as f:
        head = f.read(64).lstrip()
    if head.startswith(b"{"):
        with open(path) as f:
            return json.load(f)
    return digest_from_xml(path)
