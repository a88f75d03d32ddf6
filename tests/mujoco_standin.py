"""TEST-ONLY closed-form stand-in for the ``mujoco`` module (MuJoCo is not installable offline).

Purpose: run the reference's UNMODIFIED ``BenchmarkPlanningEnv`` / ``BenchmarkPushingEnv`` ``reset()`` / ``step()``
(envs/basic_envs.py:1770-1950 and the env callbacks) so that the ORCHESTRATION of the step path — loop order, break
semantics, noise draw sites, observation assembly, reset checks, reward / termination / info — is compared with the oracle
against the reference's own code, not against a restatement ("what's missing" #2).

What this module is: a model of exactly the MuJoCo features the two benchmark envs configure, with ``mj_step`` replaced by
the closed form the reference's own tests assert against real MuJoCo (tests/test_benchmark_planning_env.py:86-93, 199-204;
tests/test_benchmark_pushing_env.py:72-79, 165-170):

    free-joint bodies with gravcomp="1" (movers, basic_envs.py:878-879): force of a <general> actuator = gain * ctrl
    (dyntype none) or gain * act with act <- act + dt*ctrl first (dyntype integrator, actearly="true",
    planning:305-311); qacc = gear^T force / mass; semi-implicit Euler  qvel += dt*qacc ; qpos += dt*qvel.
    A body without gravcomp (the pushed object, pushing:332-342) rests on the tiles: it is kept AT REST, and the stand-in
    RAISES ``ContactError`` as soon as a mover's box geom overlaps its box geom — contact dynamics are exactly what this
    module cannot stand in for (pushing contact stays "parity unpinned").

What it is not: MuJoCo.  No constraint solver, no rotation integration (a non-zero angular velocity raises), no rendering.
It parses the XML the reference generates (names, ids, address tables, option/timestep) so that ``mujoco_utils`` — name
tables, ``mj_name2id``, ``model.actuator(name)``, ``jnt_qposadr`` ... — runs unmodified too.

Only ``tests/`` and ``tests/golden/make_golden.py`` import this file.
"""

from __future__ import annotations

import enum
import types
import xml.etree.ElementTree as ET

import numpy as np


class ContactError(RuntimeError):
    """A mover geom overlaps a resting body's geom: real MuJoCo would generate contact forces here."""


class mjtObj(enum.IntEnum):
    mjOBJ_UNKNOWN = 0
    mjOBJ_BODY = 1
    mjOBJ_XBODY = 2
    mjOBJ_JOINT = 3
    mjOBJ_DOF = 4
    mjOBJ_GEOM = 5
    mjOBJ_SITE = 6
    mjOBJ_CAMERA = 7
    mjOBJ_LIGHT = 8
    mjOBJ_FLEX = 9
    mjOBJ_MESH = 10
    mjOBJ_SKIN = 11
    mjOBJ_HFIELD = 12
    mjOBJ_TEXTURE = 13
    mjOBJ_MATERIAL = 14
    mjOBJ_PAIR = 15
    mjOBJ_EXCLUDE = 16
    mjOBJ_EQUALITY = 17
    mjOBJ_TENDON = 18
    mjOBJ_ACTUATOR = 19
    mjOBJ_SENSOR = 20


class mjtJoint(enum.IntEnum):
    mjJNT_FREE = 0
    mjJNT_BALL = 1
    mjJNT_SLIDE = 2
    mjJNT_HINGE = 3


class mjtGeom(enum.IntEnum):
    mjGEOM_PLANE = 0
    mjGEOM_HFIELD = 1
    mjGEOM_SPHERE = 2
    mjGEOM_CAPSULE = 3
    mjGEOM_ELLIPSOID = 4
    mjGEOM_CYLINDER = 5
    mjGEOM_BOX = 6
    mjGEOM_MESH = 7


def _floats(s, n=None, default=None):
    if s is None:
        return None if default is None else np.array(default, dtype=np.float64)
    v = np.array([float(x) for x in s.split()], dtype=np.float64)
    if n is not None and v.size < n:
        v = np.concatenate([v, np.zeros(n - v.size)])
    return v


class _Named:
    """What ``model.actuator(name)`` / ``model.joint(name)`` / ``model.body(id)`` return: attribute bag with ``id``."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


class MjModel:
    @staticmethod
    def from_xml_string(xml: str) -> 'MjModel':
        return MjModel(xml)

    def __init__(self, xml: str):
        root = ET.fromstring(xml.encode('utf-8') if xml.lstrip().startswith('<?xml') else xml)
        opt = root.find('option')
        self.opt = types.SimpleNamespace(
            timestep=float(opt.get('timestep', '0.002')) if opt is not None else 0.002,
            gravity=_floats(opt.get('gravity') if opt is not None else None, default=[0.0, 0.0, -9.81]),
        )
        # defaults: only geom defaults of (nested) classes matter here (tile geoms take type/size/mass from class "tile")
        self._geom_defaults: dict[str, dict[str, str]] = {}
        dflt = root.find('default')
        if dflt is not None:
            self._collect_defaults(dflt, {})

        self._names: dict[int, list[str]] = {int(t): [] for t in mjtObj}
        # bodies (id 0 = world)
        self.body_pos, self.body_gravcomp, self.body_mass_l, self.body_dofadr_l, self.body_dofnum_l = [], [], [], [], []
        self.body_parent = []
        self.jnt_type_l, self.jnt_qposadr_l, self.jnt_dofadr_l, self.jnt_bodyid_l, self.jnt_damping_l = [], [], [], [], []
        self.geom_type_l, self.geom_size_l, self.geom_bodyid_l, self.geom_pos_l = [], [], [], []
        self.nq = self.nv = 0
        wb = root.find('worldbody')
        self._add_body(None, 'world', np.zeros(3), 0.0, -1)
        if wb is not None:
            self._walk(wb, 0, None)
        # assets: meshes (names only)
        for asset in root.findall('asset'):
            for m in asset.findall('mesh'):
                self._names[int(mjtObj.mjOBJ_MESH)].append(m.get('name', ''))
        # actuators
        self.act_joint, self.act_gear, self.act_dyn, self.act_gain, self.act_early = [], [], [], [], []
        for sect in root.findall('actuator'):
            for a in sect:
                if a.tag not in ('general', 'motor'):
                    raise NotImplementedError(f'actuator <{a.tag}> is outside this stand-in')
                if a.get('biastype', 'none') != 'none':
                    raise NotImplementedError('only biastype="none"')
                self._names[int(mjtObj.mjOBJ_ACTUATOR)].append(a.get('name', ''))
                jname = a.get('joint')
                self.act_joint.append(self._names[int(mjtObj.mjOBJ_JOINT)].index(jname))
                self.act_gear.append(_floats(a.get('gear'), 6, default=[1, 0, 0, 0, 0, 0])[:6])
                self.act_dyn.append(a.get('dyntype', 'none'))
                self.act_gain.append(_floats(a.get('gainprm'), default=[1.0])[0])
                self.act_early.append(a.get('actearly', 'false') == 'true')
                if self.act_dyn[-1] not in ('none', 'integrator'):
                    raise NotImplementedError(f'dyntype {self.act_dyn[-1]}')
        for sect in root.findall('sensor'):
            for s in sect:
                self._names[int(mjtObj.mjOBJ_SENSOR)].append(s.get('name', ''))
        self.nu = len(self.act_joint)
        # activation states in actuator order
        self.act_adr = []
        na = 0
        for d in self.act_dyn:
            self.act_adr.append(na if d == 'integrator' else -1)
            na += d == 'integrator'
        self.na = na
        # array views the reference reads
        self.jnt_type = np.array(self.jnt_type_l, dtype=np.int32)
        self.jnt_qposadr = np.array(self.jnt_qposadr_l, dtype=np.int32)
        self.jnt_dofadr = np.array(self.jnt_dofadr_l, dtype=np.int32)
        self.jnt_bodyid = np.array(self.jnt_bodyid_l, dtype=np.int32)
        self.geom_type = np.array(self.geom_type_l, dtype=np.int32)
        self.geom_size = np.array(self.geom_size_l, dtype=np.float64).reshape(-1, 3)
        self.geom_bodyid = np.array(self.geom_bodyid_l, dtype=np.int32)
        self.body_mass = np.array(self.body_mass_l, dtype=np.float64)
        self.nbody, self.njnt = len(self.body_pos), len(self.jnt_type_l)
        self.ngeom = len(self.geom_type_l)
        self.nsite = len(self._names[int(mjtObj.mjOBJ_SITE)])
        self.ncam = len(self._names[int(mjtObj.mjOBJ_CAMERA)])
        self.nlight = len(self._names[int(mjtObj.mjOBJ_LIGHT)])
        self.nmesh = len(self._names[int(mjtObj.mjOBJ_MESH)])
        self.nsensor = len(self._names[int(mjtObj.mjOBJ_SENSOR)])
        self.ntendon = 0
        # MuJoCo's name buffer: NUL-terminated names, one address table per object type
        buf = bytearray()

        def table(t):
            adr = []
            for n in self._names[int(t)]:
                adr.append(len(buf))
                buf.extend(n.encode() + b'\x00')
            return np.array(adr, dtype=np.int32)

        self.name_bodyadr = table(mjtObj.mjOBJ_BODY)
        self.name_jntadr = table(mjtObj.mjOBJ_JOINT)
        self.name_geomadr = table(mjtObj.mjOBJ_GEOM)
        self.name_siteadr = table(mjtObj.mjOBJ_SITE)
        self.name_camadr = table(mjtObj.mjOBJ_CAMERA)
        self.name_lightadr = table(mjtObj.mjOBJ_LIGHT)
        self.name_meshadr = table(mjtObj.mjOBJ_MESH)
        self.name_actuatoradr = table(mjtObj.mjOBJ_ACTUATOR)
        self.name_sensoradr = table(mjtObj.mjOBJ_SENSOR)
        self.name_tendonadr = np.zeros(0, dtype=np.int32)
        self.names = bytes(buf)

    # ---- XML walk
    def _collect_defaults(self, node, inherited):
        cur = dict(inherited)
        g = node.find('geom')
        if g is not None:
            cur.update(g.attrib)
        cls = node.get('class')
        if cls is not None:
            self._geom_defaults[cls] = cur
        for child in node.findall('default'):
            self._collect_defaults(child, cur)

    def _add_body(self, parent, name, pos, gravcomp, parent_id):
        self._names[int(mjtObj.mjOBJ_BODY)].append(name)
        self.body_pos.append(np.asarray(pos, dtype=np.float64))
        self.body_gravcomp.append(float(gravcomp))
        self.body_mass_l.append(0.0)
        self.body_dofadr_l.append(-1)
        self.body_dofnum_l.append(0)
        self.body_parent.append(parent_id)
        return len(self.body_pos) - 1

    def _walk(self, node, body_id, childclass):
        for el in node:
            if el.tag == 'body':
                if body_id != 0:
                    raise NotImplementedError('nested bodies are outside this stand-in')
                bid = self._add_body(node, el.get('name', ''), _floats(el.get('pos'), 3, default=[0, 0, 0]),
                                     float(el.get('gravcomp', '0')), body_id)
                self._walk(el, bid, el.get('childclass', childclass))
            elif el.tag in ('joint', 'freejoint'):
                jt = 'free' if el.tag == 'freejoint' else el.get('type', 'hinge')
                if jt != 'free':
                    raise NotImplementedError('only free joints')
                self._names[int(mjtObj.mjOBJ_JOINT)].append(el.get('name', ''))
                self.jnt_type_l.append(int(mjtJoint.mjJNT_FREE))
                self.jnt_qposadr_l.append(self.nq)
                self.jnt_dofadr_l.append(self.nv)
                self.jnt_bodyid_l.append(body_id)
                self.jnt_damping_l.append(float(el.get('damping', '0')))
                self.body_dofadr_l[body_id] = self.nv
                self.body_dofnum_l[body_id] = 6
                self.nq += 7
                self.nv += 6
            elif el.tag == 'geom':
                at = dict(self._geom_defaults.get(el.get('class', childclass), {})) if (el.get('class') or childclass) else {}
                at.update(el.attrib)
                self._names[int(mjtObj.mjOBJ_GEOM)].append(at.get('name', ''))
                gt = {'plane': mjtGeom.mjGEOM_PLANE, 'sphere': mjtGeom.mjGEOM_SPHERE, 'capsule': mjtGeom.mjGEOM_CAPSULE,
                      'ellipsoid': mjtGeom.mjGEOM_ELLIPSOID, 'cylinder': mjtGeom.mjGEOM_CYLINDER, 'box': mjtGeom.mjGEOM_BOX,
                      'mesh': mjtGeom.mjGEOM_MESH}[at.get('type', 'sphere')]
                self.geom_type_l.append(int(gt))
                self.geom_size_l.append(_floats(at.get('size'), 3, default=[0, 0, 0])[:3])
                self.geom_pos_l.append(_floats(at.get('pos'), 3, default=[0, 0, 0]))
                self.geom_bodyid_l.append(body_id)
                if at.get('mass') is not None:
                    self.body_mass_l[body_id] += float(at['mass'])
            elif el.tag == 'site':
                self._names[int(mjtObj.mjOBJ_SITE)].append(el.get('name', ''))
            elif el.tag == 'camera':
                self._names[int(mjtObj.mjOBJ_CAMERA)].append(el.get('name', ''))
            elif el.tag == 'light':
                self._names[int(mjtObj.mjOBJ_LIGHT)].append(el.get('name', ''))

    # ---- named access (the subset the reference uses)
    def _id(self, objtype, key) -> int:
        if isinstance(key, (int, np.integer)):
            return int(key)
        try:
            return self._names[int(objtype)].index(key)
        except ValueError:
            raise KeyError(f'no {mjtObj(objtype).name} named {key!r}') from None

    def actuator(self, key):
        return _Named(id=self._id(mjtObj.mjOBJ_ACTUATOR, key))

    def joint(self, key):
        i = self._id(mjtObj.mjOBJ_JOINT, key)
        return _Named(id=i, bodyid=np.array([self.jnt_bodyid_l[i]]), qposadr=np.array([self.jnt_qposadr_l[i]]),
                      dofadr=np.array([self.jnt_dofadr_l[i]]))

    def body(self, key):
        i = self._id(mjtObj.mjOBJ_BODY, key)
        return _Named(id=i, dofadr=np.array([self.body_dofadr_l[i]]), dofnum=np.array([self.body_dofnum_l[i]]),
                      mass=np.array([self.body_mass_l[i]]))


class MjData:
    def __init__(self, model: MjModel):
        self.qpos = np.zeros(model.nq)
        self.qvel = np.zeros(model.nv)
        self.qacc = np.zeros(model.nv)
        self.ctrl = np.zeros(model.nu)
        self.act = np.zeros(model.na)
        self.time = 0.0
        self.xpos = np.zeros((model.nbody, 3))
        self.xmat = np.tile(np.eye(3).reshape(1, 9), (model.nbody, 1))
        for j in range(model.njnt):
            a, b = model.jnt_qposadr_l[j], model.jnt_bodyid_l[j]
            self.qpos[a:a + 3] = model.body_pos[b]
            self.qpos[a + 3:a + 7] = (1.0, 0.0, 0.0, 0.0)
        _kinematics(model, self)


def _kinematics(model: MjModel, data: MjData) -> None:
    for b in range(model.nbody):
        data.xpos[b] = model.body_pos[b]
    for j in range(model.njnt):
        a, b = model.jnt_qposadr_l[j], model.jnt_bodyid_l[j]
        data.xpos[b] = data.qpos[a:a + 3]
        q = data.qpos[a + 3:a + 7]
        if not (q[0] == 1.0 and q[1] == 0.0 and q[2] == 0.0 and q[3] == 0.0):
            raise NotImplementedError('the stand-in keeps every body at the identity orientation')


def _forward(model: MjModel, data: MjData, advance_act: bool) -> np.ndarray:
    """qacc of the closed form.  Returns the activation vector the step would commit (act + dt*ctrl)."""
    dt = model.opt.timestep
    act_next = data.act.copy()
    force = np.zeros(model.nv)
    for i in range(model.nu):
        j = model.act_joint[i]
        if model.act_dyn[i] == 'integrator':
            k = model.act_adr[i]
            act_next[k] = data.act[k] + dt * data.ctrl[i]  # act_dot = ctrl (Euler)
            drive = act_next[k] if model.act_early[i] else data.act[k]
        else:
            drive = data.ctrl[i]
        mass = model.body_mass_l[model.jnt_bodyid_l[j]]
        # gain == mass for the movers' x/y actuators (planning:305-321): the ratio is exactly 1 and qacc == drive
        d0 = model.jnt_dofadr_l[j]
        gear = model.act_gear[i]
        if np.any(gear[3:] != 0.0):
            if drive != 0.0:
                raise NotImplementedError('rotational actuation is outside the closed form (impedance torque must be zero)')
            continue
        force[d0:d0 + 3] += gear[:3] * ((model.act_gain[i] / mass) * drive)  # force / mass
    qacc = np.zeros(model.nv)
    for j in range(model.njnt):
        b, d0 = model.jnt_bodyid_l[j], model.jnt_dofadr_l[j]
        if model.body_gravcomp[b] == 1.0:
            qacc[d0:d0 + 3] = force[d0:d0 + 3]  # gravity exactly compensated
        else:
            # a body under gravity rests on the tiles (normal force balances gravity): at rest unless something touches it
            if np.any(force[d0:d0 + 6] != 0.0) or np.any(data.qvel[d0:d0 + 6] != 0.0):
                raise NotImplementedError('a moving non-gravity-compensated body is outside the closed form')
    data.qacc[:] = qacc
    return act_next


def _check_contacts(model: MjModel, data: MjData) -> None:
    """Raise ContactError if a gravity-compensated body's box geom overlaps a resting body's box geom (axis-aligned)."""
    boxes = {}
    for g in range(model.ngeom):
        b = int(model.geom_bodyid_l[g])
        if model.body_dofnum_l[b] == 0 or model.geom_type_l[g] != int(mjtGeom.mjGEOM_BOX):
            continue
        boxes.setdefault(b, []).append((data.xpos[b] + model.geom_pos_l[g], model.geom_size_l[g]))
    ids = sorted(boxes)
    for i, bi in enumerate(ids):
        for bj in ids[i + 1:]:
            if model.body_gravcomp[bi] == 1.0 and model.body_gravcomp[bj] == 1.0:
                continue  # mover-mover: the envs flag the collision before the geoms touch
            for ci, si in boxes[bi]:
                for cj, sj in boxes[bj]:
                    if np.all(np.abs(ci - cj) <= si + sj):
                        raise ContactError(f'bodies {bi} and {bj} touch: contact dynamics are not modelled')


def mj_forward(model: MjModel, data: MjData) -> None:
    _kinematics(model, data)
    _forward(model, data, advance_act=False)


def mj_step(model: MjModel, data: MjData, nstep: int = 1) -> None:
    dt = model.opt.timestep
    for _ in range(nstep):
        _kinematics(model, data)
        _check_contacts(model, data)
        act_next = _forward(model, data, advance_act=True)
        data.act[:] = act_next
        for j in range(model.njnt):
            a, d0 = model.jnt_qposadr_l[j], model.jnt_dofadr_l[j]
            if np.any(data.qvel[d0 + 3:d0 + 6] != 0.0):
                raise NotImplementedError('angular velocity: rotation integration is outside the closed form')
            # semi-implicit Euler (tests/test_benchmark_planning_env.py:86-93): qvel += dt*qacc ; qpos += dt*qvel
            data.qvel[d0:d0 + 3] = data.qvel[d0:d0 + 3] + dt * data.qacc[d0:d0 + 3]
            data.qpos[a:a + 3] = data.qpos[a:a + 3] + dt * data.qvel[d0:d0 + 3]
        data.time += dt
    _kinematics(model, data)


def mj_name2id(model: MjModel, objtype, name: str) -> int:
    try:
        return model._names[int(objtype)].index(name)
    except ValueError:
        return -1


def mj_jacBody(model: MjModel, data: MjData, jacp, jacr, body_id: int) -> None:
    """Jacobian of a free-joint body's frame origin at the identity orientation: translation dofs -> I, rotation -> I."""
    if jacp is not None:
        jacp[:] = 0.0
    if jacr is not None:
        jacr[:] = 0.0
    d0 = model.body_dofadr_l[int(body_id)]
    if d0 < 0:
        return
    if jacp is not None:
        jacp[:, d0:d0 + 3] = np.eye(3)
    if jacr is not None:
        jacr[:, d0 + 3:d0 + 6] = np.eye(3)


def as_module() -> types.ModuleType:
    """The object to put into ``sys.modules['mujoco']``."""
    m = types.ModuleType('mujoco')
    m.__dict__.update(MjModel=MjModel, MjData=MjData, mjtObj=mjtObj, mjtJoint=mjtJoint, mjtGeom=mjtGeom, mj_step=mj_step,
                      mj_forward=mj_forward, mj_name2id=mj_name2id, mj_jacBody=mj_jacBody, ContactError=ContactError,
                      __standin__=True)
    viewer = types.ModuleType('mujoco.viewer')
    m.viewer = viewer
    return m
