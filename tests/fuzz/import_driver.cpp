#include "opr.h"
#include <cstdio>
int main(int argc, char** argv) {
    void* h = opr_scene_file_load(argv[1], argc > 2 ? 1 : 0);
    if (!h) { printf("err %.80s\n", opr_scene_file_last_error()); return 0; }
    const opr_scene_desc* d = opr_scene_file_desc(h);
    // touch everything the renderer would read
    double acc = 0; long cnt = 0;
    for (int m = 0; m < d->numMeshes; m++) {
        const opr_mesh& me = d->meshes[m];
        for (int t = 0; t < me.numTriangles * 3; t++) { int i = me.indices[t]; acc += me.vertices[3 * i] + (me.normals ? me.normals[3 * i] : 0) + (me.uvs ? me.uvs[2 * i] : 0); cnt++; }
    }
    for (int i = 0; i < d->numLights; i++) acc += d->lights[i].power[0];
    for (int i = 0; i < d->numMaterials; i++) acc += d->materials[i].kind;
    opr_camera c; opr_scene_file_camera(h, &c);
    printf("ok %ld %g %s\n", cnt, acc, opr_scene_file_warnings(h) ? "w" : "");
    opr_scene_file_destroy(h);
    return 0;
}
