"""Writes a small COLLADA 1.4 room (test fixture): floor with a TGA texture, walls, an emissive ceiling quad, a glass box, a
mirror panel, a pentagon without normals, a point light and a camera, placed through translate / rotate / scale / matrix nodes."""
import os
import struct

import numpy as np

QUAD_POS = [(-1, 0, -1), (1, 0, -1), (1, 0, 1), (-1, 0, 1)]          # y-up unit quad of side 2, normal +y (CCW seen from +y: 0 3 2 1)
QUAD_UV = [(0, 0), (1, 0), (1, 1), (0, 1)]
BOX = [(-1, -1, -1), (1, -1, -1), (1, 1, -1), (-1, 1, -1), (-1, -1, 1), (1, -1, 1), (1, 1, 1), (-1, 1, 1)]
BOX_TRIS = [(0, 2, 1), (0, 3, 2), (4, 5, 6), (4, 6, 7), (0, 1, 5), (0, 5, 4), (2, 3, 7), (2, 7, 6), (1, 2, 6), (1, 6, 5), (0, 4, 7), (0, 7, 3)]
PENTAGON = [(np.cos(2 * np.pi * k / 5), 0.0, -np.sin(2 * np.pi * k / 5)) for k in range(5)]


def _fa(name, rows):
    flat = " ".join("%.9g" % v for r in rows for v in r)
    n = len(rows) * len(rows[0])
    return ('<source id="%s"><float_array id="%s-a" count="%d">%s</float_array><technique_common><accessor source="#%s-a" count="%d" stride="%d"/>'
            '</technique_common></source>' % (name, name, n, flat, name, len(rows), len(rows[0])))


def effect(eid, emission=None, diffuse=None, reflective=None, ior=None, texture=None, bump=None):
    params, slots = "", ""
    if texture:
        params += ('<newparam sid="%s-surf"><surface type="2D"><init_from>%s</init_from></surface></newparam>'
                   '<newparam sid="%s-samp"><sampler2D><source>%s-surf</source></sampler2D></newparam>' % (eid, texture, eid, eid))
    if emission:
        slots += "<emission><color>%g %g %g 1</color></emission>" % emission
    if texture:
        slots += '<diffuse><texture texture="%s-samp" texcoord="UV0"/></diffuse>' % eid
    elif diffuse:
        slots += "<diffuse><color>%g %g %g 1</color></diffuse>" % diffuse
    if reflective:
        slots += "<reflective><color>%g %g %g 1</color></reflective>" % reflective
    if ior:
        slots += "<index_of_refraction><float>%g</float></index_of_refraction>" % ior
    extra = ""
    if bump:
        params += ('<newparam sid="%s-bsurf"><surface type="2D"><init_from>%s</init_from></surface></newparam>'
                   '<newparam sid="%s-bsamp"><sampler2D><source>%s-bsurf</source></sampler2D></newparam>' % (eid, bump, eid, eid))
        extra = '<extra><technique profile="FCOLLADA"><bump><texture texture="%s-bsamp" texcoord="UV0"/></bump></technique></extra>' % eid
    return '<effect id="%s"><profile_COMMON>%s<technique sid="common"><phong>%s</phong>%s</technique></profile_COMMON></effect>' % (eid, params, slots, extra)


def write_tga(path, pixels_bgr):
    h, w, d = pixels_bgr.shape
    open(path, "wb").write(bytes([0, 0, 2]) + bytes(5) + struct.pack("<HHHH", 0, 0, w, h) + bytes([8 * d, 0]) + pixels_bgr.tobytes())


def write_room(dirname, up_axis="Y_UP"):
    os.makedirs(dirname, exist_ok=True)
    rng = np.random.default_rng(11)
    tex = (rng.integers(60, 200, (8, 8, 3))).astype(np.uint8)
    write_tga(os.path.join(dirname, "floor.tga"), tex)
    nrm = np.zeros((8, 8, 3), np.uint8)
    nrm[...] = (255, 128, 128)          # BGR of the flat normal (0.5, 0.5, 1)
    write_tga(os.path.join(dirname, "floor_n.tga"), nrm)
    effects = [effect("fx-white", diffuse=(0.8, 0.8, 0.8)), effect("fx-red", diffuse=(0.8, 0.1, 0.1)), effect("fx-light", emission=(40, 40, 36), diffuse=(1, 1, 1)),
               effect("fx-glass", diffuse=(1, 1, 1), ior=1.5), effect("fx-mirror", diffuse=(0, 0, 0), reflective=(0.9, 0.9, 0.9)),
               effect("fx-floor", texture="img-floor", bump="img-floor-n")]
    names = ["white", "red", "light", "glass", "mirror", "floor"]
    mats = "".join('<material id="mat-%s" name="%s"><instance_effect url="#fx-%s"/></material>' % (n, n.upper(), n) for n in names)
    quad = ('<geometry id="geo-quad"><mesh>%s%s%s<vertices id="quad-v"><input semantic="POSITION" source="#quad-p"/></vertices>'
            '<polylist material="M" count="1"><input semantic="VERTEX" source="#quad-v" offset="0"/><input semantic="NORMAL" source="#quad-n" offset="1"/>'
            '<input semantic="TEXCOORD" source="#quad-t" offset="2" set="0"/><vcount>4</vcount><p>0 0 0 3 0 3 2 0 2 1 0 1</p></polylist></mesh></geometry>'
            % (_fa("quad-p", QUAD_POS), _fa("quad-n", [(0, 1, 0)]), _fa("quad-t", QUAD_UV)))
    box = ('<geometry id="geo-box"><mesh>%s<vertices id="box-v"><input semantic="POSITION" source="#box-p"/></vertices>'
           '<triangles material="M" count="12"><input semantic="VERTEX" source="#box-v" offset="0"/><p>%s</p></triangles></mesh></geometry>'
           % (_fa("box-p", BOX), " ".join("%d %d %d" % t for t in BOX_TRIS)))
    pent = ('<geometry id="geo-pent"><mesh>%s<vertices id="pent-v"><input semantic="POSITION" source="#pent-p"/></vertices>'
            '<polygons material="M" count="1"><input semantic="VERTEX" source="#pent-v" offset="0"/><p>0 1 2 3 4</p></polygons></mesh></geometry>' % _fa("pent-p", PENTAGON))

    def inst(geo, mat):
        return ('<instance_geometry url="#geo-%s"><bind_material><technique_common><instance_material symbol="M" target="#mat-%s"/></technique_common>'
                '</bind_material></instance_geometry>' % (geo, mat))
    nodes = [
        '<node id="n-floor"><scale>2 1 2</scale>%s</node>' % inst("quad", "floor"),
        '<node id="n-ceil"><translate>0 4 0</translate><rotate>1 0 0 180</rotate><scale>2 1 2</scale>%s</node>' % inst("quad", "white"),
        '<node id="n-back"><translate>0 2 2</translate><rotate>1 0 0 -90</rotate><scale>2 1 2</scale>%s</node>' % inst("quad", "white"),
        '<node id="n-left"><matrix>0 1 0 -2  -2 0 0 2  0 0 2 0  0 0 0 1</matrix>%s</node>' % inst("quad", "red"),
        '<node id="n-right"><translate>2 2 0</translate><rotate>0 0 1 90</rotate><scale>2 1 2</scale>%s</node>' % inst("quad", "white"),
        '<node id="n-light"><translate>0 3.98 0</translate><rotate>1 0 0 180</rotate><scale>0.5 1 0.5</scale>%s</node>' % inst("quad", "light"),
        '<node id="n-glass"><translate>-0.8 0.61 0.3</translate><rotate>0 1 0 30</rotate><scale>0.4 0.6 0.4</scale>%s</node>' % inst("box", "glass"),
        '<node id="n-mirror"><translate>1.95 1.6 0.4</translate><rotate>0 0 1 90</rotate><scale>0.8 1 0.6</scale>%s</node>' % inst("quad", "mirror"),
        '<node id="n-group"><translate>0.8 0 -0.5</translate><node id="n-pent"><translate>0 0.3 0</translate><scale>0.5 1 0.5</scale>%s</node></node>' % inst("pent", "red"),
        '<node id="n-plight"><translate>1.2 3.0 -1.0</translate><instance_light url="#light-point"/></node>',
        '<node id="n-cam"><translate>0 2 -5.5</translate><rotate>0 1 0 180</rotate><instance_camera url="#cam"/></node>',
    ]
    doc = ('<?xml version="1.0" encoding="utf-8"?>\n<!-- test fixture -->\n<COLLADA xmlns="http://www.collada.org/2005/11/COLLADASchema" version="1.4.1">'
           '<asset><unit name="meter" meter="1"/><up_axis>%s</up_axis></asset>'
           '<library_images><image id="img-floor"><init_from>floor.tga</init_from></image><image id="img-floor-n"><init_from>./floor_n.tga</init_from></image></library_images>'
           '<library_effects>%s</library_effects><library_materials>%s</library_materials>'
           '<library_geometries>%s%s%s</library_geometries>'
           '<library_lights><light id="light-point"><technique_common><point><color>30 30 28</color></point></technique_common></light></library_lights>'
           '<library_cameras><camera id="cam"><optics><technique_common><perspective><xfov>50</xfov><aspect_ratio>1</aspect_ratio><znear>0.1</znear><zfar>100</zfar>'
           '</perspective></technique_common></optics></camera></library_cameras>'
           '<library_visual_scenes><visual_scene id="vs">%s</visual_scene></library_visual_scenes><scene><instance_visual_scene url="#vs"/></scene></COLLADA>'
           % (up_axis, "".join(effects), mats, quad, box, pent, "".join(nodes)))
    path = os.path.join(dirname, "room.dae")
    open(path, "w").write(doc)
    return path, tex
