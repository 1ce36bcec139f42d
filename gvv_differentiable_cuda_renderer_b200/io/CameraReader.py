"""Skeletool '.calibration' reader -- mirror of python/utils/CameraReader.py.

Same attributes as the reference (python/utils/CameraReader.py:5-48): `extrinsics` (flat list,
12 per camera: the first three rows of the 4x4), `intrinsics` (flat list, 9 per camera: the 3x3
part of the 4x4, rescaled from the calibration's sensor size to the render resolution, :41-46),
`numberOfCameras`, `originalSizeU/V`.
"""
import numpy as np


class CameraReader:
    def __init__(self, filename, renderResolutionU, renderResolutionV):
        self.filename = filename
        self.extrinsics, self.intrinsics = [], []
        self.originalSizeU, self.originalSizeV = [], []
        with open(filename) as fh:
            for line in fh:
                tok = line.split()
                if not tok:
                    continue
                if tok[0] == 'intrinsic':
                    vals = [float(t) for t in tok[1:17]]
                    self.intrinsics += [vals[4 * r + c] for r in range(3) for c in range(3)]
                elif tok[0] == 'extrinsic':
                    self.extrinsics += [float(t) for t in tok[1:13]]
                elif tok[0] == 'size':
                    self.originalSizeU.append(float(tok[1]))
                    self.originalSizeV.append(float(tok[2]))
        self.numberOfCameras = len(self.extrinsics) // 12
        K = np.asarray(self.intrinsics, dtype=np.float64).reshape(self.numberOfCameras, 3, 3)
        for c in range(self.numberOfCameras):
            # same operation order as the reference (divide by the calibration size, then multiply, :41-46), in float64
            K[c, 0, 0] = (K[c, 0, 0] / self.originalSizeU[c]) * renderResolutionU
            K[c, 1, 1] = (K[c, 1, 1] / self.originalSizeV[c]) * renderResolutionV
            K[c, 0, 2] = (K[c, 0, 2] / self.originalSizeU[c]) * renderResolutionU
            K[c, 1, 2] = (K[c, 1, 2] / self.originalSizeV[c]) * renderResolutionV
        self.intrinsics = list(K.flatten())

    def extrinsics_array(self):
        return np.asarray(self.extrinsics, np.float32).reshape(1, -1)

    def intrinsics_array(self):
        return np.asarray(self.intrinsics, np.float32).reshape(1, -1)
