// stub for declarations in RayTracer/AssetManager.h:81-88 (the loader itself is not compiled)
#pragma once
struct aiNode; struct aiScene; struct aiMesh;
