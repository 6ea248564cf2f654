"""Known-answer tests for the interactive camera: Camera::Pitch / Yaw / Roll / Move* of include/venusaur/Camera.h
against the reference's formulas (camera.h:40-78), which are glm quaternion code: rotation = glm::rotate(quat(1,0,0,0), radians(speed),
axis); forward = normalize(forward * rotation).  glm is not in this image, so the test restates the three glm functions involved from
their published definitions (glm/gtc/quaternion.inl, glm/detail/type_quat.inl, version-independent since 0.9.5) in float32:
    rotate(q, angle, v):  v is normalised when |length(v) - 1| > 0.001;  q * quat(cos(angle/2), v * sin(angle/2))
    v * q              :  inverse(q) * v,  inverse(q) = conjugate(q) / dot(q, q)
    q * v              :  uv = cross(q.xyz, v); uuv = cross(q.xyz, uv); v + ((uv * q.w) + uuv) * 2
The drop-in header uses Rodrigues' formula instead (same rotation, different rounding): agreement to a few float ulps is required."""
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
f32 = np.float32


def _normalize(v):
    return (v * (f32(1.0) / np.sqrt(np.dot(v, v), dtype=f32))).astype(f32)


def _glm_rotate_identity(angle, axis):
    axis = axis.astype(f32)
    ln = np.sqrt(np.dot(axis, axis), dtype=f32)
    if abs(ln - f32(1.0)) > f32(0.001):
        axis = (axis * (f32(1.0) / ln)).astype(f32)
    s = f32(np.sin(f32(angle) * f32(0.5), dtype=f32))
    c = f32(np.cos(f32(angle) * f32(0.5), dtype=f32))
    return c, (axis * s).astype(f32)            # quat(1,0,0,0) * q = q


def _glm_vec_times_quat(v, w, q):
    d = f32(w * w + np.dot(q, q))
    iw, iq = f32(w / d), (-q / d).astype(f32)   # inverse(q)
    uv = np.cross(iq, v).astype(f32)
    uuv = np.cross(iq, uv).astype(f32)
    return (v + ((uv * iw) + uuv) * f32(2.0)).astype(f32)


class GlmCamera:
    """camera.h:9-100 + Camera.cpp:16-37, literally, in float32."""

    def __init__(self, pos, vfov, aspect, aperture, focal):
        self.pos, self.vfov, self.aspect, self.focal = np.array(pos, f32), f32(vfov), f32(aspect), f32(focal)
        self.forward = np.array([0, 0, -1], f32)
        self.update()

    def update(self):
        self.w = (self.forward * self.focal).astype(f32)
        self.u = _normalize(np.cross(self.w, np.array([0, 1, 0], f32)).astype(f32))
        self.v = _normalize(np.cross(self.u, self.w).astype(f32))
        h = f32(np.tan(f32(np.radians(self.vfov, dtype=f32)) * f32(0.5), dtype=f32))
        vh = f32(2.0) * h
        vw = self.aspect * vh
        self.u = (self.u * (self.focal * vw)).astype(f32)
        self.v = (self.v * (self.focal * vh)).astype(f32)

    def rotate(self, speed, axis):
        w, q = _glm_rotate_identity(np.radians(f32(speed), dtype=f32), axis)
        self.forward = _normalize(_glm_vec_times_quat(self.forward, w, q))
        # (no update(): the reference only refreshes m_u, m_v, m_w in UVWFrame(), once per frame -- several key events between two frames
        #  all rotate about the axes of the last frame, and so does the drop-in class)

    def frame(self):
        self.update()           # Camera::UVWFrame, Camera.cpp:16-22
        return np.concatenate([self.u, self.v, self.w, self.pos])


SRC = r'''
#include <cstdio>
#include "venusaur/Camera.h"
using venusaur::vec3;
static void dump(venusaur::Camera& c) {
    vec3 u, v, w; c.UVWFrame(u, v, w);
    const vec3 p = c.GetPosition();
    printf("%.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g\n", u.x, u.y, u.z, v.x, v.y, v.z, w.x, w.y, w.z, p.x, p.y, p.z);
}
int main() {
    venusaur::Camera c(vec3(13, 2, 3), 20.0f, 1.5f, 0.1f, 10.0f);
    c.SetForward(vec3(-13, -2, -3));
    dump(c);
    c.Pitch(0.5f); dump(c);
    c.Yaw(-0.5f); dump(c);
    c.Roll(3.0f); dump(c);
    for (int i = 0; i < 40; i++) { c.Yaw(0.5f); c.Pitch(-0.25f); }
    dump(c);
    c.MoveForward(0.1f); c.MoveRight(-0.2f); c.MoveUp(0.3f); dump(c);
    printf("%d\n", (int)c.Changed());
    printf("%d\n", (int)c.Changed());
    return 0;
}'''


def test_camera_rotation_and_motion_match_the_glm_formulas(tmp_path):
    cpp, exe = tmp_path / "cam.cpp", tmp_path / "cam"
    cpp.write_text(SRC)
    subprocess.run(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-I" + os.path.join(ROOT, "include"), str(cpp), "-o", str(exe)], check=True)
    lines = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.strip().splitlines()
    got = [np.array([float(x) for x in ln.split()], f32) for ln in lines[:6]]
    assert lines[6:] == ["1", "0"]                               # Changed() reports and clears the dirty flag (camera.h:80-85)

    ref = GlmCamera((13, 2, 3), 20.0, 1.5, 0.1, 10.0)
    ref.forward = _normalize(np.array([-13, -2, -3], f32))      # SetForward, camera.h:24
    ref.update()
    want = [ref.frame()]
    ref.rotate(0.5, ref.u); want.append(ref.frame())            # Pitch: about m_u (camera.h:56-62)
    ref.rotate(-0.5, ref.v); want.append(ref.frame())           # Yaw: about m_v (:64-70)
    ref.rotate(3.0, ref.w); want.append(ref.frame())            # Roll: about m_w (:72-78) -- forward itself: no change
    for _ in range(40):                                         # 80 key events within one frame: the axes stay those of the last frame
        ref.rotate(0.5, ref.v)
        ref.rotate(-0.25, ref.u)
    want.append(ref.frame())
    ref.pos = (ref.pos + ref.forward * f32(0.1)).astype(f32)                                     # MoveForward (:40-45)
    ref.pos = (ref.pos + _normalize(ref.u) * f32(-0.2)).astype(f32)                              # MoveRight (:47-51)
    ref.pos = (ref.pos + _normalize(ref.v) * f32(0.3)).astype(f32)                               # MoveUp (:53-58)
    want.append(ref.frame())
    for k, (g, w) in enumerate(zip(got, want)):
        scale = max(1.0, float(np.abs(w).max()))
        assert np.abs(g - w).max() <= 4e-6 * scale, (k, g, w)
    # Roll about the view direction leaves the forward vector (hence the whole frame) where it was: the reference's Roll is a no-op too
    assert np.abs(got[3] - got[2]).max() <= 2e-6 * 13.0
    # 40 x (0.5 deg yaw, -0.25 deg pitch) really moved the view
    assert np.abs(got[4][6:9] - got[3][6:9]).max() > 1.0
