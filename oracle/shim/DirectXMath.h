/*
 * oracle/shim/DirectXMath.h -- oracle build only.
 * The reference uses exactly one DirectXMath call on the path: XMMatrixMultiply in
 * mat44f::mult (Core/Math.cpp:120-124).  This is a scalar restatement of the published
 * row-major 4x4 product with the SSE path's summation order ((x*r0 + z*r2) + (y*r1 + w*r3)).
 */
#pragma once
namespace DirectX {
struct XMFLOAT4X4 { float _11,_12,_13,_14,_21,_22,_23,_24,_31,_32,_33,_34,_41,_42,_43,_44; };
struct XMVECTOR { float v[4]; };
struct XMMATRIX {
    float m[4][4];
    XMMATRIX() {}
    explicit XMMATRIX(const float* p) { for (int i = 0; i < 16; ++i) (&m[0][0])[i] = p[i]; }
};
inline XMMATRIX XMMatrixMultiply(const XMMATRIX& a, const XMMATRIX& b) {
    XMMATRIX r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            float x = a.m[i][0] * b.m[0][j];
            float y = a.m[i][1] * b.m[1][j];
            float z = a.m[i][2] * b.m[2][j];
            float w = a.m[i][3] * b.m[3][j];
            r.m[i][j] = (x + z) + (y + w);
        }
    return r;
}
inline void XMStoreFloat4x4(XMFLOAT4X4* dst, const XMMATRIX& s) {
    float* d = &dst->_11; for (int i = 0; i < 16; ++i) d[i] = (&s.m[0][0])[i];
}
}
