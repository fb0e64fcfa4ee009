// rtc_camera.cpp -- host-side camera math of the path (no GPU involved).
// Restates Camera3D::Init / Update / GetInverseVMatrix and the block Engine3D::Render
// assembles (reference Camera3D.cpp:8-48, :51-98, :207-376; Engine3D.cpp:88-97) in binary32,
// same operations in the same order (compile without FMA contraction).
#include <cmath>
#include <cstdint>

#include "rtc_kernels.h"

namespace rtc {

int camera_params(uint32_t x, uint32_t y, const float pos[3], const float rot[3], float pixel_aspect, rtc_params* out)
{
    if (!out || !pos || !rot || x == 0 || y == 0) return RTC_ERR_INVALID;
    const float k = pixel_aspect == 0.0f ? 0.01f : pixel_aspect;       // Camera3D.cpp:17
    const float fov = (float)(M_PI) / 1.5f;                            // Camera3D.h:79, .cpp:10
    const float width = (float)x, height = (float)y;
    const float aspect = width / (k * width * height);
    const float e = 1.0f / std::tan(fov / 2.0f);                       // :19
    const float p = rot[0], yw = rot[1];
    const float fwd[3] = {-std::sin(yw), -std::sin(p) * std::cos(yw), -std::cos(p) * std::cos(yw)};    // :57-59
    const float right[3] = {std::cos(yw), -std::sin(p) * std::sin(yw), -std::cos(p) * std::sin(yw)};   // :65-67
    const float up[3] = {0.0f, std::cos(p), -std::sin(p)};                                              // :73-75
    const float m[4][4] = {{right[0], up[0], fwd[0], pos[0]},          // :79-98: basis vectors as COLUMNS
                           {right[1], up[1], fwd[1], pos[1]},
                           {right[2], up[2], fwd[2], pos[2]},
                           {0.0f, 0.0f, 0.0f, 1.0f}};
    // Cofactor inverse (:210-343): entry (r,c) is the signed 3x3 minor deleting row c / column r,
    // expanded in the reference's fixed six-term order; odd entries carry the flipped signs term by
    // term (this matters only for the sign of an exactly-zero entry).
    float inv[4][4];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
            int R[3], C[3], n = 0;
            for (int i = 0; i < 4; ++i) if (i != c) R[n++] = i;
            n = 0;
            for (int i = 0; i < 4; ++i) if (i != r) C[n++] = i;
            const float t1 = m[R[0]][C[0]] * m[R[1]][C[1]] * m[R[2]][C[2]];
            const float t2 = m[R[0]][C[0]] * m[R[1]][C[2]] * m[R[2]][C[1]];
            const float t3 = m[R[1]][C[0]] * m[R[0]][C[1]] * m[R[2]][C[2]];
            const float t4 = m[R[1]][C[0]] * m[R[0]][C[2]] * m[R[2]][C[1]];
            const float t5 = m[R[2]][C[0]] * m[R[0]][C[1]] * m[R[1]][C[2]];
            const float t6 = m[R[2]][C[0]] * m[R[0]][C[2]] * m[R[1]][C[1]];
            inv[r][c] = ((r + c) & 1) ? (-t1 + t2 + t3 - t4 - t5 + t6) : (t1 - t2 - t3 + t4 + t5 - t6);
        }
    float det = m[0][0] * inv[0][0] + m[0][1] * inv[1][0] + m[0][2] * inv[2][0] + m[0][3] * inv[3][0];   // :345-349
    if (det == 0.0f) return RTC_ERR_INVALID;                           // reference asserts (:351)
    det = 1.0f / det;
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) out->inv_view[4 * r + c] = inv[r][c] * det;
    out->cam_pos[0] = pos[0]; out->cam_pos[1] = pos[1]; out->cam_pos[2] = pos[2];
    out->x = x; out->y = y;
    out->element1 = e / aspect;    // m_pMatrix.row1.x (:29)
    out->element2 = e;             // m_pMatrix.row2.y (:35)
    out->cam_far = 250.0f;         // Camera3D.h:75
    return RTC_OK;
}

}  // namespace rtc
