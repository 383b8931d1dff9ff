/*
 * opr_scene.h — plain-C data descriptions that cross the drop-in boundary: Camera, Light, materials,
 * geometry and the request details.  No CUDA, torch or C++ types.
 *
 * These mirror the reference's PODs byte for byte where the reference uploads them as user data:
 *   opr_camera  == class Camera      (RenderEngine/renderer/Camera.h:72-79, 96 bytes)
 *   opr_light   == class Light       (RenderEngine/renderer/Light.h:27-53, 88 bytes)
 * and replace the two types that cannot exist without OptiX/Qt:
 *   IScene::getSceneRootGroup(optix::Context&) (RenderEngine/scene/IScene.h:21)  -> opr_scene_desc
 *   QVector<Light>                                                              -> (lights, numLights)
 * The material record carries exactly the values the reference passes through
 * registerGeometryInstanceValues (material/Diffuse.cpp:48-51, DiffuseEmitter.cpp:48-64,
 * Glass.cpp:58-64, Mirror.cpp:42-45, Glossy.cpp:61-66, Texture.cpp:102-107).
 */
#ifndef OPR_SCENE_H
#define OPR_SCENE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* RenderEngine/renderer/RenderMethod.h:13-19 (same order, same values) */
enum opr_render_method {
    OPR_PATH_TRACING = 0,
    OPR_VCM_BIDIRECTIONAL_PATH_TRACING = 1,
    OPR_PROGRESSIVE_PHOTON_MAPPING = 2,
    OPR_RENDER_METHOD_NONE = 3
};

/* Camera::AspectRatioMode, RenderEngine/renderer/Camera.h:18-23 */
enum opr_aspect_ratio_mode { OPR_KEEP_VERTICAL = 0, OPR_KEEP_HORIZONTAL = 1, OPR_KEEP_NONE = 2 };

/* Light::LightType, RenderEngine/renderer/Light.h:17 */
enum opr_light_type { OPR_LIGHT_AREA = 0, OPR_LIGHT_POINT = 1, OPR_LIGHT_SPOT = 2 };

/* one OptiX Material singleton per class in the reference (material/<Kind>.cpp); here a kind id */
enum opr_material_kind {
    OPR_MAT_DIFFUSE = 0,
    OPR_MAT_DIFFUSE_EMITTER = 1,
    OPR_MAT_GLASS = 2,
    OPR_MAT_MIRROR = 3,
    OPR_MAT_TEXTURE = 4,
    OPR_MAT_GLOSSY = 5,
    OPR_MAT_NUM_KINDS = 6
};

typedef struct opr_camera {
    float eye[3], lookat[3], up[3];
    float hfov, vfov;
    float aperture;
    float lookdir[3];      /* not normalised: its length is the focal distance */
    float camera_u[3], camera_v[3];
    float imagePlaneSize[2];
    int32_t aspectRatioMode;
} opr_camera;               /* 96 bytes */

typedef struct opr_light {
    float power[3];
    float position[3];
    float v1[3];
    float v2[3];
    uint8_t isDelta;
    uint8_t isFinite;
    uint8_t _pad[2];
    float inverseArea;
    float area_or_angle;    /* union { area; angle } */
    float Lemit[3];         /* union { Lemit; intensity } */
    float normal[3];        /* union { normal; direction } */
    int32_t lightType;
} opr_light;                /* 88 bytes */

typedef struct opr_material {
    int32_t kind;           /* opr_material_kind */
    float Kd[3];            /* Diffuse, Glossy, DiffuseEmitter */
    float Ks[3];            /* Glossy */
    float exponent;         /* Glossy */
    float Kr[3];            /* Mirror, Glass */
    float Kt[3];            /* Glass */
    float indexOfRefraction;/* Glass */
    float power[3];         /* DiffuseEmitter: m_power (already multiplied by Kd, DiffuseEmitter.cpp:16-21) */
    float inverseArea;      /* DiffuseEmitter */
    int32_t diffuseTexture; /* Texture: index into opr_scene_desc.textures, else -1 */
    int32_t normalTexture;  /* Texture: index or -1 (hasNormals) */
} opr_material;

/* RGBA8, row-major, sampled with wrap=repeat, bilinear, normalised float reads, no mips
 * (material/Texture.cpp:109-121) */
typedef struct opr_texture {
    int32_t width, height;
    const uint8_t* rgba;
} opr_texture;

/* One triangle mesh = one material (scene/Scene.cpp:496-503).  normals/uvs/tangents/bitangents may be
 * NULL (geometry_instance/TriangleMesh.cu:49-88). */
typedef struct opr_mesh {
    int32_t numVertices;
    int32_t numTriangles;
    const float* vertices;   /* 3*numVertices */
    const float* normals;    /* 3*numVertices or NULL */
    const float* uvs;        /* 2*numVertices or NULL */
    const float* tangents;   /* 3*numVertices or NULL */
    const float* bitangents; /* 3*numVertices or NULL */
    const int32_t* indices;  /* 3*numTriangles */
    int32_t material;
} opr_mesh;

/* arguments of Cornell::createParallelogram (scene/Cornell.cpp:33-64): un-scaled offsets */
typedef struct opr_parallelogram {
    float anchor[3], offset1[3], offset2[3];
    int32_t material;
} opr_parallelogram;

/* SphereInstance (geometry_instance/SphereInstance.cpp:35-36) */
typedef struct opr_sphere {
    float center[3];
    float radius;
    int32_t material;
} opr_sphere;

typedef struct opr_scene_desc {
    const char* name;
    int32_t numMeshes;         const opr_mesh* meshes;
    int32_t numParallelograms; const opr_parallelogram* parallelograms;
    int32_t numSpheres;        const opr_sphere* spheres;
    int32_t numMaterials;      const opr_material* materials;
    int32_t numTextures;       const opr_texture* textures;
    int32_t numLights;         const opr_light* lights;
    float aabbMin[3], aabbMax[3];   /* IScene::getSceneAABB() */
} opr_scene_desc;

/* RenderServerRenderRequestDetails getters (clientserver/RenderServerRenderRequestDetails.h:15-33) */
typedef struct opr_request_details {
    opr_camera camera;
    int32_t renderMethod;   /* opr_render_method */
    uint32_t width, height;
    float ppmAlpha;
    const char* sceneName;  /* may be NULL */
} opr_request_details;

/* Runtime knobs for values the reference hard-codes (SURVEY.md §5 "Config / flags"):
 * OptixRenderer.cpp:44-53, config.h:33-37, config.h:50.  Zero = reference default. */
typedef struct opr_config {
    uint32_t photonLaunchWidth;   /* default 1024 */
    uint32_t photonLaunchHeight;  /* default 1024 */
    uint32_t vcmMaxPathLength;    /* default 10 */
    uint32_t randomSeed;          /* default 1645301512 (config.h:50 DEBUG_RANDOM_SEED) */
    uint32_t seedIsSet;           /* 0 => use default seed above */
} opr_config;

#ifdef __cplusplus
}
#endif
#endif
