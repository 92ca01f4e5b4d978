"""Camera-parameter loading from hardware_parameters/*.txt -- Python twin of the reference's
CParameterBase (src/utility/CParameterBase.h:21-66 tokeniser, :88-141 getters, :169-226
loadCameraLEFT/RIGHT).  Same file format, same error behaviour (missing key -> ParameterError,
malformed number -> ValueError like std::invalid_argument)."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


class ParameterError(Exception):
    """CExceptionParameter (src/exceptions/CExceptionParameter.h)."""


def get_parameters_from_file(path: str) -> list[str]:
    """getParametersFromFile :21-66: every non-empty line split on single spaces, flattened."""
    try:
        f = open(path, "r")
    except OSError:
        raise ParameterError("unable to open file: '%s'" % path)
    tokens: list[str] = []
    with f:
        for line in f.read().split("\n"):
            if line:
                tokens.extend(line.split(" "))
    return tokens


def _find(tokens, name):
    try:
        return tokens.index(name)
    except ValueError:
        raise ParameterError("cannot find parameter: " + name)


def get_double(tokens, name) -> float:
    return float(tokens[_find(tokens, name) + 1])


def get_integer(tokens, name) -> int:
    v = int(tokens[_find(tokens, name) + 1])
    if v < 0:
        raise ValueError("negative value for " + name)
    return v


def get_matrix(tokens, name, rows, cols) -> np.ndarray:
    i = _find(tokens, name)
    vals = [float(t) for t in tokens[i + 1:i + 1 + rows * cols]]
    if len(vals) != rows * cols:
        raise ValueError("not enough values for " + name)
    return np.asarray(vals, np.float64).reshape(rows, cols)


@dataclass
class PinholeCamera:
    """The CPinholeCamera members the hot path uses (src/vision/CPinholeCamera.h:16-64,202-227)."""
    label: str
    width: int
    height: int
    P: np.ndarray                 # m_matProjection 3x4
    K: np.ndarray                 # m_matIntrinsic 3x3
    focal_length_m: float
    distortion: np.ndarray
    rectification: np.ndarray

    @property
    def fov(self):
        """m_cFieldOfView = Rect(28, 28, W-56, H-56) (CPinholeCamera.h:61) as (x, y, w, h)."""
        return (28, 28, self.width - 56, self.height - 56)

    def principal_weight_u(self, u: float) -> float:
        return float(np.sqrt(abs(u - self.P[0, 2])) / 10.0)

    def principal_weight_v(self, v: float) -> float:
        return float(np.sqrt(abs(v - self.P[1, 2])) / 10.0)


def load_camera(path: str) -> PinholeCamera:
    """loadCameraLEFT / loadCameraRIGHT :169-226 (identical bodies)."""
    t = get_parameters_from_file(path)
    if not t:
        raise ParameterError("unable to open file: '%s'" % path)
    return PinholeCamera(
        label=t[0],
        width=get_integer(t, "uWidthPixels"),
        height=get_integer(t, "uHeightPixels"),
        P=get_matrix(t, "matProjection", 3, 4),
        K=get_matrix(t, "matIntrinsic", 3, 3),
        focal_length_m=get_double(t, "dFocalLengthMeters"),
        distortion=get_matrix(t, "vecDistortionCoefficients", 4, 1).reshape(4),
        rectification=get_matrix(t, "matRectification", 3, 3),
    )


@dataclass
class StereoCamera:
    """CStereoCamera (src/vision/CStereoCamera.h:14-35): the pair plus the manual baseline vector."""
    left: PinholeCamera
    right: PinholeCamera
    translation_to_right: np.ndarray

    @property
    def baseline_m(self) -> float:
        return float(np.linalg.norm(self.translation_to_right))


def construct_camera_stereo(left: PinholeCamera, right: PinholeCamera, translation_to_right=(-0.54, 0.0, 0.0)) -> StereoCamera:
    """constructCameraSTEREO(Vector3d) :312-318 (tracker_gt.cpp:123 passes (-0.54, 0, 0))."""
    return StereoCamera(left, right, np.asarray(translation_to_right, np.float64))
