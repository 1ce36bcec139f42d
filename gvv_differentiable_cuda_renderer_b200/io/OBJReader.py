"""Wavefront OBJ reader for the renderer's inputs -- mirror of python/utils/OBJReader.py.

Same attribute names as the reference class (python/utils/OBJReader.py:41-90,104-196), so the
call sites of its scripts keep working:
    facesVertexId, facesTextureId         flat int lists, 3 per face (0-based)
    vertexCoordinates, vertexColors       [N][3]   ("v x y z r g b" lines, the reference's dialect)
    pertVertexTextureCoordinate           [T][2]   ("vt u v")
    textureCoordinates                    flat float list, 6 per face (per-corner u,v) = the op attribute
    numberOfVertices, textureMap (H x W x 3 in [0,1], RGB), texHeight, texWidth
    compressedAdjacency, numberOfNeigbours, numberOfEdges, maximumNumNeighbours

Differences, on purpose: quads and n-gons keep their first three corners like the reference, but
the adjacency is stored sparsely only (the reference also builds dense N x N float matrices,
105 MB at 5k vertices and impossible at 35k, python/utils/OBJReader.py:108); vertex colours default
to 0.5 when the file has none; a missing MTL/texture is not an error (textureMap = None).
"""
import os

import numpy as np


def _load_image_rgb01(path):
    try:
        from PIL import Image
        # uint8 / 255.0 in float64 like the reference (cv2 image / 255.0, OBJReader.py:203-205), then fp32 as the op casts it
        return (np.asarray(Image.open(path).convert("RGB"), dtype=np.float64) / 255.0).astype(np.float32)
    except ImportError:
        import cv2
        img = cv2.imread(path)
        return (cv2.cvtColor(img, cv2.COLOR_BGR2RGB) / 255.0).astype(np.float32)


class OBJReader:
    def __init__(self, filename, verbose=False):
        self.filename = filename
        self.folderPath = filename[0:filename.rfind('/') + 1]
        self.mtlFilePath = self.folderPath
        self.mtlFilePathFull = None
        self.readObjFile()
        self.numberOfVertices = len(self.vertexColors)
        self.computePerFaceTextureCoordinated()
        self.loadSegmentationWeights()
        self.computeAdjacency()
        self.textureMap, self.texHeight, self.texWidth = None, 0, 0
        if self.mtlFilePathFull is not None and os.path.exists(self.mtlFilePathFull):
            self.loadMtlTexture(self.mtlFilePathFull, self.mtlFilePath)
        if verbose:
            print(f'++ ObjReader: {self.numberOfVertices} vertices, {len(self.facesVertexId) // 3} faces')

    def readObjFile(self):
        self.facesVertexId, self.facesTextureId = [], []
        self.vertexColors, self.vertexCoordinates, self.pertVertexTextureCoordinate = [], [], []
        with open(self.filename) as fh:
            for line in fh:
                tok = line.split()
                if not tok:
                    continue
                if tok[0] == 'f':
                    for corner in tok[1:4]:
                        idx = corner.split('/')
                        self.facesVertexId.append(int(idx[0]) - 1)
                        self.facesTextureId.append(int(idx[1]) - 1 if len(idx) > 1 and idx[1] else -1)
                elif tok[0] == 'v':
                    self.vertexCoordinates.append([float(tok[1]), float(tok[2]), float(tok[3])])
                    self.vertexColors.append([float(tok[4]), float(tok[5]), float(tok[6])] if len(tok) >= 7 else [0.5, 0.5, 0.5])
                elif tok[0] == 'vt':
                    self.pertVertexTextureCoordinate.append([float(tok[1]), float(tok[2])])
                elif tok[0] == 'mtllib':
                    name = tok[1][2:] if tok[1].startswith('./') else tok[1]
                    self.mtlFilePathFull = self.folderPath + name

    def computePerFaceTextureCoordinated(self):
        self.textureCoordinates = []
        for t in self.facesTextureId:
            u, v = self.pertVertexTextureCoordinate[t] if t >= 0 else (0.0, 0.0)
            self.textureCoordinates.append(u)
            self.textureCoordinates.append(v)

    def computeAdjacency(self):
        nb = [set() for _ in range(self.numberOfVertices)]
        f = self.facesVertexId
        for i in range(0, len(f), 3):
            a, b, c = f[i], f[i + 1], f[i + 2]
            nb[a].update((b, c)); nb[b].update((a, c)); nb[c].update((a, b))
        for v, s in enumerate(nb):
            s.discard(v)
        self.compressedAdjacency = [sorted(s) for s in nb]
        self.numberOfNeigbours = np.asarray([len(s) for s in nb], dtype=np.float32)
        self.numberOfEdges = int(self.numberOfNeigbours.sum())
        self.maximumNumNeighbours = int(self.numberOfNeigbours.max()) if self.numberOfVertices else 0

    def loadMtlTexture(self, mtlFileName, shortPath):
        with open(mtlFileName) as fh:
            for line in fh:
                tok = line.split()
                if tok and tok[0] == 'map_Kd':
                    path = shortPath + tok[1]
                    if os.path.exists(path):
                        self.textureMap = _load_image_rgb01(path)
                        self.texHeight, self.texWidth = self.textureMap.shape[0], self.textureMap.shape[1]

    def loadSegmentationWeights(self):
        """Per-vertex labels of 'segmentation.txt' next to the OBJ, when present (:99-...)."""
        self.vertexLabels = []
        path = self.folderPath + 'segmentation.txt'
        if os.path.exists(path):
            with open(path) as fh:
                self.vertexLabels = [int(l.split()[0]) for l in fh if l.split()]

    # convenience for the op
    def faces_array(self):
        return np.asarray(self.facesVertexId, np.int32).reshape(-1, 3)

    def texcoords_array(self):
        return np.asarray(self.textureCoordinates, np.float32).reshape(-1, 3, 2)
