// nao.cuh -- the Nao-cup scenario: state validity and the validator of its edge check.
//
// Replaces nao_clear / nao_link of the reference (demo/nao_cup/src/naocup.hpp:556-730,795-840; primitives
// demo/nao_cup/src/collide.hpp:45-115, linear.hpp:125-150; scenario demo/nao_cup_planning.cpp:146-152):
// forward kinematics of two 5-joint arms, 209 sphere / capsule pair tests (the reference evaluates every one, :592-613;
// the answer is their OR), the cup must stay within 25 degrees of upright (:687-689).
//
// The reference multiplies general 4x4 isometries.  Every rotation there is about a coordinate axis and every
// translation has one or two non-zero entries, so most products are with exact zeros and ones.  This file uses the
// closed forms that remain when those are removed -- x*0 = +-0, s + (+-0) = s, x*1 = x, x + (-y) = x - y are exact, and
// the sign of a zero never reaches a decision -- so every value below equals the one the general product yields (with
// Eigen's operation structure, sums left to right, unfused): same decisions, about a fifth of the arithmetic.  The CPU
// oracle (oracle/oracle_nao.hpp) keeps the general form, so the parity tests check exactly this claim.
//   AngleAxis about x: [[u,0,0],[0,c,-s],[0,s,c]], about y: [[c,0,s],[0,u,0],[-s,0,c]], about z: [[c,-s,0],[s,c,0],[0,0,u]]
//   with u = (1 - c) + c as Eigen's toRotationMatrix forms the diagonal (NOT always 1).
// Distances: the reference compares sqrt(x) < R.  sqrt is correctly rounded and monotone, so that is x < T(R) with
// T(R) = the smallest x whose rounded root reaches R, found once per radius sum on the host: no square root and no
// branch per pair; the three cases of the segment-point distance (linear.hpp:133-149) become two selects.
// Tried and dropped: skipping a capsule test when the centres are farther apart than the half lengths plus the reach (a
// conservative bound with a 1/32 margin, decisions unchanged: the parity tests passed).  It removes most of the 107 capsule
// tests of a configuration on paper, but a branch per pair costs more than the tests it saves: lanes of a warp hold midpoints
// of different edges and disagree, and the straight-line version keeps the FP32 pipe fed -- 65,536 edges 0.3 rad apart
// 0.68 -> 1.12 ms (float), states 2.39 -> 2.18 G/s.
#pragma once

#include <cmath>

#include "../../include/mptg/mptg_fpmath.h"

namespace mptg {
namespace nao {

constexpr int DIM = 10;
constexpr int BEADS = 8;

template <typename S>
struct Capsule {  // segment p0-p1 with v = p1 - p0 and c2 = v.v as linear.hpp:129,138 form them
    S p0[3], p1[3], v[3], c2;
};

// Everything that does not depend on the state, evaluated once on the host in the scalar type of the geometry.
template <typename S>
struct Model {
    // joints: constant rotations by +-pi/2 (naocup.hpp:384,390,403)
    S sH, cH, sN, cN, uN;
    S shoulderY, shoulderZ, upperArm;
    S handX, handZ, handOff;  // :400-405
    S armLen;                 // lower-arm capsule, :430-431
    S ballOff;                // :573
    S stemZ, bowlZ, stemLen, bowlLen;  // :574-577, :527-535
    S beadX[BEADS], beadY[BEADS], beadZTop, beadZBottom;  // :540-543, :579-587
    S bowlCentreZ, upGoal;    // :676-677, :689
    S disc;                   // DISCRETIZATION, :80
    // static objects (:426-520, transforms :567-571 and compute_head :351-372)
    S torso[3][3];
    S headC[3];
    Capsule<S> headCap, coke, pepsi;
    S table[3], wall[3];
    // thresholds: T(ra + rb) for tests that take a root, (ra + rb)^2 for sphere-sphere (collide.hpp:86-90)
    S tTorsoArm[3], tArmArm, tHeadArm, tEarArm, tArmCoke, tArmPepsi, tPlaneArm;
    S tTorsoStem[3], tTorsoBowl[3], r2TorsoBead[3];
    S tStemCoke, tStemPepsi, tBowlCoke, tBowlPepsi, tPlaneStem, tPlaneBowl, tBeadCoke, tBeadPepsi, r2PlaneBead;
    S tStemArm, tBowlArm, tBeadArm;
    S tHeadStem, tHeadBowl, tStemEar, tBowlEar, r2HeadBead, tBeadEar;
    S tBallArm, r2HeadBall, tBallEar, r2TorsoBall[3], tBallCoke, tBallPepsi, r2PlaneBall;
};

// smallest x with fl(sqrt(x)) >= R  (R > 0)
template <typename S>
inline S rootThreshold(S R) {
    S x = R * R;
    while (x > S(0) && fp::sqrt_(x) >= R) x = std::nextafter(x, S(-1));
    while (fp::sqrt_(x) < R) x = std::nextafter(x, fp::consts<S>::inf());
    return x;
}
template <typename S>
inline S sumSquared(S a, S b) {  // collide.hpp:87-88
    S r = a + b;
    r = r * r;
    return r;
}
template <typename S>
inline void setCapsule(Capsule<S>& c, const S p0[3], const S p1[3]) {
    for (int i = 0; i < 3; ++i) c.p0[i] = p0[i], c.p1[i] = p1[i], c.v[i] = p1[i] - p0[i];
    c.c2 = (c.v[0] * c.v[0] + c.v[1] * c.v[1]) + c.v[2] * c.v[2];
}

template <typename S>
inline Model<S> makeModel() {
    Model<S> m;
    // naocup.hpp:55-80,222-251 -- each constant with the expression types of the reference
    const S inch = S(0.0254), pi = S(3.14159265358979323846);
    const S centerTorsoR = S(66.7 / 1000.0), beadR = S(2.0 / 4.0) * inch, cupDiam = S(2.5) * inch;
    const S cupHeight = (S(4.0) + S(3.0) / S(8.0)) * inch, gripHeight = S(1.0) * inch;
    const S baseToBowl = (S(1.0) + S(5.0) / S(8.0)) * inch, gripDiam = S(5.0) / S(8.0) * inch;
    const S gripCapsuleHeight = baseToBowl - gripDiam * S(2.0), bowlHeight = cupHeight - baseToBowl;
    const S ballR = S(0.015), planeR = S(25.0), tableZ = S(0.09);
    const S neckZ = S(126.50 / 1000.0), lowerArm = S(50.55 / 1000.0), handOffX = S(58.00 / 1000.0), hipZ = S(85.00 / 1000.0);
    const S handOffZ = S(15.90 / 1000.0), headR = S(115.0 / 2.0 / 1000.0), earR = S(90.0 / 2.0 / 1000.0), headW = S(133.0 / 1000.0);
    const S armR = S(66.7 / 2.0 / 1000.0), handR = S(20.0 / 1000.0), handW = S(50.0 / 1000.0);
    fp::sincos_(pi / S(2.0), &m.sH, &m.cH);
    fp::sincos_(-pi / S(2.0), &m.sN, &m.cN);
    m.uN = (S(1) - m.cN) + m.cN;
    m.shoulderY = S(98.00 / 1000.0), m.shoulderZ = S(100.00 / 1000.0), m.upperArm = S(90.00 / 1000.0);
    m.handX = lowerArm + handOffX + S(0.01);
    m.handZ = -handOffZ - S(0.01);
    m.handOff = (handW - handR * S(2)) / S(2);
    m.armLen = lowerArm + handOffX - armR;
    m.ballOff = ballR / S(2.0) + S(0.01);
    m.stemZ = gripHeight / S(2.0);
    m.bowlZ = gripHeight / S(2.0) + cupDiam / S(2.0);
    m.stemLen = -gripCapsuleHeight;
    m.bowlLen = bowlHeight - cupDiam;
    for (int i = 0; i < BEADS; ++i) {  // :540-543: angle in S, ::cos / ::sin(double), product in double
        const S a = pi * S(2.0) * (S)i / (S)BEADS;
        double sn, cs;
        fp::sincos_((double)a, &sn, &cs);
        const S r = cupDiam / S(2.0) - beadR;
        m.beadX[i] = (S)(cs * (double)r);
        m.beadY[i] = (S)(sn * (double)r);
    }
    m.beadZTop = bowlHeight + gripHeight / S(2.0) - beadR;
    m.beadZBottom = gripHeight / S(2.0) - baseToBowl + beadR;
    m.bowlCentreZ = S(bowlHeight + gripHeight / S(2.0));
    m.upGoal = S(0.90630778703665);
    m.disc = S(1.0) * pi / S(180.0);

    // robot frame = identity: a translation by (x,y,z) lands exactly on (x,y,z)
    const S torsoR[3] = {S(55.6 / 1000.0), centerTorsoR, hipZ / S(2.0)};
    const S torsoZ[3] = {S(0), neckZ - S(66.7) / S(1000.0), -hipZ / S(2.0)};
    for (int k = 0; k < 3; ++k) m.torso[k][0] = S(0), m.torso[k][1] = S(0), m.torso[k][2] = torsoZ[k];
    // head: yaw = pitch = 0 (fields of the value-initialised world that nothing sets), so both rotations are the
    // identity exactly (sin 0 = 0, cos 0 = 1, u = (1-1)+1 = 1); centre = (0, 0, neckZ + headR);
    // head capsule frame = centre * Rx(pi/2) * Translation(0,0,z): linear part [[u,0,0],[0,c,-s],[0,s,c]]
    m.headC[0] = S(0), m.headC[1] = S(0), m.headC[2] = neckZ + headR;
    {
        const S z = -(headW - earR * S(2)) / S(2), len = headW - earR / S(2.0);
        const S col2[3] = {S(0), -m.sH, m.cH};  // third column of Rx(pi/2)
        S p0[3], p1[3];
        for (int i = 0; i < 3; ++i) p0[i] = m.headC[i] + col2[i] * z;
        for (int i = 0; i < 3; ++i) p1[i] = col2[i] * len + p0[i];
        setCapsule(m.headCap, p0, p1);
    }
    const S cokeR = S(2.5) / S(2.0) * inch, cokeLen = (S(6.75) - S(2.5) / S(2.0)) * inch;
    const S pepsiR = S(3.0) / S(2.0) * inch, pepsiLen = (S(8.5) - S(3.0) / S(2.0)) * inch;
    {
        const S p0[3] = {S(0.12), S(0.08), -tableZ};
        const S p1[3] = {p0[0], p0[1], cokeLen + p0[2]};
        setCapsule(m.coke, p0, p1);
        const S q0[3] = {S(0.12) + S(2.5) * inch, S(-0.12), -tableZ};
        const S q1[3] = {q0[0], q0[1], pepsiLen + q0[2]};
        setCapsule(m.pepsi, q0, q1);
    }
    m.table[0] = S(0), m.table[1] = S(0), m.table[2] = -planeR - tableZ;
    m.wall[0] = -planeR - centerTorsoR, m.wall[1] = S(0), m.wall[2] = S(0);

    const S stemR = gripDiam / S(2.0), bowlR = cupDiam / S(2.0);
    for (int k = 0; k < 3; ++k) {
        m.tTorsoArm[k] = rootThreshold<S>(torsoR[k] + armR);
        m.tTorsoStem[k] = rootThreshold<S>(torsoR[k] + stemR);
        m.tTorsoBowl[k] = rootThreshold<S>(torsoR[k] + bowlR);
        m.r2TorsoBead[k] = sumSquared<S>(beadR, torsoR[k]);
        m.r2TorsoBall[k] = sumSquared<S>(torsoR[k], ballR);
    }
    m.tArmArm = rootThreshold<S>(armR + armR);
    m.tHeadArm = rootThreshold<S>(headR + armR);
    m.tEarArm = rootThreshold<S>(earR + armR);
    m.tArmCoke = rootThreshold<S>(armR + cokeR);
    m.tArmPepsi = rootThreshold<S>(armR + pepsiR);
    m.tPlaneArm = rootThreshold<S>(planeR + armR);
    m.tStemCoke = rootThreshold<S>(stemR + cokeR), m.tStemPepsi = rootThreshold<S>(stemR + pepsiR);
    m.tBowlCoke = rootThreshold<S>(bowlR + cokeR), m.tBowlPepsi = rootThreshold<S>(bowlR + pepsiR);
    m.tPlaneStem = rootThreshold<S>(planeR + stemR), m.tPlaneBowl = rootThreshold<S>(planeR + bowlR);
    m.tBeadCoke = rootThreshold<S>(beadR + cokeR), m.tBeadPepsi = rootThreshold<S>(beadR + pepsiR);
    m.r2PlaneBead = sumSquared<S>(beadR, planeR);
    m.tStemArm = rootThreshold<S>(stemR + armR), m.tBowlArm = rootThreshold<S>(bowlR + armR), m.tBeadArm = rootThreshold<S>(beadR + armR);
    m.tHeadStem = rootThreshold<S>(headR + stemR), m.tHeadBowl = rootThreshold<S>(headR + bowlR);
    m.tStemEar = rootThreshold<S>(stemR + earR), m.tBowlEar = rootThreshold<S>(bowlR + earR);
    m.r2HeadBead = sumSquared<S>(beadR, headR);
    m.tBeadEar = rootThreshold<S>(beadR + earR);
    m.tBallArm = rootThreshold<S>(ballR + armR);
    m.r2HeadBall = sumSquared<S>(headR, ballR);
    m.tBallEar = rootThreshold<S>(ballR + earR);
    m.tBallCoke = rootThreshold<S>(ballR + cokeR), m.tBallPepsi = rootThreshold<S>(ballR + pepsiR);
    m.r2PlaneBall = sumSquared<S>(planeR, ballR);
    return m;
}

#ifdef __CUDACC__
// L <- L * AngleAxis(about x / y / z) with (s, c, u) of the angle; rows left to right as the general product
template <typename S>
__device__ __forceinline__ void rotX(S L[3][3], S s, S c, S u) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const S a = L[i][1], b = L[i][2];
        L[i][0] = L[i][0] * u;
        L[i][1] = a * c + b * s;
        L[i][2] = b * c - a * s;
    }
}
template <typename S>
__device__ __forceinline__ void rotZ(S L[3][3], S s, S c, S u) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const S a = L[i][0], b = L[i][1];
        L[i][0] = a * c + b * s;
        L[i][1] = b * c - a * s;
        L[i][2] = L[i][2] * u;
    }
}

// squared distance argument of v3_dist_segment_point (linear.hpp:125-150) before its root
template <typename S>
__device__ __forceinline__ S segPointArg(const S p0[3], const S p1[3], const S v[3], S c2, const S pt[3]) {
    const S w0 = pt[0] - p0[0], w1 = pt[1] - p0[1], w2 = pt[2] - p0[2];
    const S c1 = (w0 * v[0] + w1 * v[1]) + w2 * v[2];
    const S ww = (w0 * w0 + w1 * w1) + w2 * w2;
    const S e0 = pt[0] - p1[0], e1 = pt[1] - p1[1], e2 = pt[2] - p1[2];
    const S ee = (e0 * e0 + e1 * e1) + e2 * e2;
    const S b = fp::div_(c1, c2);
    const S f0 = (p0[0] + v[0] * b) - pt[0], f1 = (p0[1] + v[1] * b) - pt[1], f2 = (p0[2] + v[2] * b) - pt[2];
    const S ff = (f0 * f0 + f1 * f1) + f2 * f2;
    return c1 <= S(0) ? ww : (c2 <= c1 ? ee : ff);
}
template <typename S>
__device__ __forceinline__ S segPointArg(const Capsule<S>& c, const S pt[3]) {
    return segPointArg<S>(c.p0, c.p1, c.v, c.c2, pt);
}
// collide_capsule_capsule (collide.hpp:93-115): the smallest of the four end-to-segment distances
template <typename S>
__device__ __forceinline__ S capCapArg(const Capsule<S>& a, const Capsule<S>& b) {
    const S a0 = segPointArg<S>(a, b.p0), a1 = segPointArg<S>(a, b.p1);
    const S b0 = segPointArg<S>(b, a.p0), b1 = segPointArg<S>(b, a.p1);
    return fmin(fmin(a0, a1), fmin(b0, b1));
}
template <typename S>
__device__ __forceinline__ S dist2(const S a[3], const S b[3]) {  // (a - b).squaredNorm(), collide.hpp:86
    const S d0 = a[0] - b[0], d1 = a[1] - b[1], d2 = a[2] - b[2];
    return (d0 * d0 + d1 * d1) + d2 * d2;
}

template <typename S>
struct Hand {  // the part of an arm the tests read
    Capsule<S> arm;   // lower-arm capsule
    S L[3][3], t[3];  // transform_to_hand
};

// compute_arm (naocup.hpp:375-406) for the base frame Translation(0, y, shoulderZ) of the identity
template <typename S>
__device__ __forceinline__ void arm(const Model<S>& m, S baseY, const S* q, Hand<S>& h) {
    S s, c;
    S L[3][3];
    fp::sincos_(q[0], &s, &c);  // shoulder pitch about y, applied to the identity: the rotation itself
    L[0][0] = c, L[0][1] = S(0), L[0][2] = s;
    L[1][0] = S(0), L[1][1] = (S(1) - c) + c, L[1][2] = S(0);
    L[2][0] = -s, L[2][1] = S(0), L[2][2] = c;
    fp::sincos_(q[1], &s, &c);  // shoulder roll about z
    rotZ<S>(L, s, c, (S(1) - c) + c);
    S t[3];  // to the elbow: + linear * (upperArm, 0, 0)
    t[0] = S(0) + L[0][0] * m.upperArm;
    t[1] = baseY + L[1][0] * m.upperArm;
    t[2] = m.shoulderZ + L[2][0] * m.upperArm;
    fp::sincos_(q[2], &s, &c);  // elbow yaw about x
    rotX<S>(L, s, c, (S(1) - c) + c);
    fp::sincos_(q[3], &s, &c);  // elbow roll about z
    rotZ<S>(L, s, c, (S(1) - c) + c);
    // lower capsule frame = elbow roll * Ry(pi/2): only its third column (the capsule axis) is read
    S p1[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) p1[i] = (L[i][0] * m.sH + L[i][2] * m.cH) * m.armLen + t[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) h.arm.p0[i] = t[i], h.arm.p1[i] = p1[i], h.arm.v[i] = p1[i] - t[i];
    h.arm.c2 = (h.arm.v[0] * h.arm.v[0] + h.arm.v[1] * h.arm.v[1]) + h.arm.v[2] * h.arm.v[2];
    fp::sincos_(q[4], &s, &c);  // wrist yaw about x
    rotX<S>(L, s, c, (S(1) - c) + c);
#pragma unroll
    for (int i = 0; i < 3; ++i) t[i] = t[i] + (L[i][0] * m.handX + L[i][2] * m.handZ);
    rotX<S>(L, m.sN, m.cN, m.uN);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        h.t[i] = t[i] + L[i][2] * m.handOff;
        h.L[i][0] = L[i][0], h.L[i][1] = L[i][1], h.L[i][2] = L[i][2];
    }
}

// nao_clear (naocup.hpp:795-806)
template <typename S>
__device__ bool clear(const Model<S>& m, const S* q) {
    Hand<S> R;
    arm<S>(m, -m.shoulderY, q, R);
    // cup_is_up (:673-689)
    {
        S up[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) up[i] = (R.L[i][2] * m.bowlCentreZ + R.t[i]) - R.t[i];
        const S n2 = (up[0] * up[0] + up[1] * up[1]) + up[2] * up[2];
        const S z = n2 > S(0) ? fp::div_(up[2], fp::sqrt_(n2)) : up[2];
        if (!(z > m.upGoal)) return false;
    }
    Hand<S> Lh;
    arm<S>(m, m.shoulderY, q + 5, Lh);
    bool hit = false;
    // arms against the body and the obstacles (:594-602)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        hit |= segPointArg<S>(R.arm, m.torso[k]) < m.tTorsoArm[k];
        hit |= segPointArg<S>(Lh.arm, m.torso[k]) < m.tTorsoArm[k];
    }
    hit |= capCapArg<S>(Lh.arm, R.arm) < m.tArmArm;
    hit |= segPointArg<S>(R.arm, m.headC) < m.tHeadArm;
    hit |= segPointArg<S>(Lh.arm, m.headC) < m.tHeadArm;
    hit |= capCapArg<S>(m.headCap, R.arm) < m.tEarArm;
    hit |= capCapArg<S>(m.headCap, Lh.arm) < m.tEarArm;
    hit |= capCapArg<S>(R.arm, m.coke) < m.tArmCoke;
    hit |= capCapArg<S>(R.arm, m.pepsi) < m.tArmPepsi;
    hit |= segPointArg<S>(R.arm, m.table) < m.tPlaneArm;
    hit |= segPointArg<S>(R.arm, m.wall) < m.tPlaneArm;
    hit |= capCapArg<S>(Lh.arm, m.coke) < m.tArmCoke;
    hit |= capCapArg<S>(Lh.arm, m.pepsi) < m.tArmPepsi;
    hit |= segPointArg<S>(Lh.arm, m.table) < m.tPlaneArm;
    hit |= segPointArg<S>(Lh.arm, m.wall) < m.tPlaneArm;
    if (hit) return false;
    // ball in the left hand (:573, :609-612)
    {
        S ball[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) ball[i] = Lh.t[i] + Lh.L[i][1] * m.ballOff;
        hit |= segPointArg<S>(R.arm, ball) < m.tBallArm;
        hit |= dist2<S>(m.headC, ball) < m.r2HeadBall;
        hit |= segPointArg<S>(m.headCap, ball) < m.tBallEar;
#pragma unroll
        for (int k = 0; k < 3; ++k) hit |= dist2<S>(m.torso[k], ball) < m.r2TorsoBall[k];
        hit |= segPointArg<S>(m.coke, ball) < m.tBallCoke;
        hit |= segPointArg<S>(m.pepsi, ball) < m.tBallPepsi;
        hit |= dist2<S>(m.table, ball) < m.r2PlaneBall;
        hit |= dist2<S>(m.wall, ball) < m.r2PlaneBall;
    }
    // cup stem and bowl in the right hand (:574-577, :603-608)
#pragma unroll
    for (int part = 0; part < 2; ++part) {
        const S z = part == 0 ? m.stemZ : m.bowlZ, len = part == 0 ? m.stemLen : m.bowlLen;
        Capsule<S> cap;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            cap.p0[i] = R.t[i] + R.L[i][2] * z;
            cap.p1[i] = R.L[i][2] * len + cap.p0[i];
            cap.v[i] = cap.p1[i] - cap.p0[i];
        }
        cap.c2 = (cap.v[0] * cap.v[0] + cap.v[1] * cap.v[1]) + cap.v[2] * cap.v[2];
#pragma unroll
        for (int k = 0; k < 3; ++k) hit |= segPointArg<S>(cap, m.torso[k]) < (part == 0 ? m.tTorsoStem[k] : m.tTorsoBowl[k]);
        hit |= capCapArg<S>(cap, m.coke) < (part == 0 ? m.tStemCoke : m.tBowlCoke);
        hit |= capCapArg<S>(cap, m.pepsi) < (part == 0 ? m.tStemPepsi : m.tBowlPepsi);
        hit |= segPointArg<S>(cap, m.table) < (part == 0 ? m.tPlaneStem : m.tPlaneBowl);
        hit |= segPointArg<S>(cap, m.wall) < (part == 0 ? m.tPlaneStem : m.tPlaneBowl);
        hit |= capCapArg<S>(cap, Lh.arm) < (part == 0 ? m.tStemArm : m.tBowlArm);
        hit |= segPointArg<S>(cap, m.headC) < (part == 0 ? m.tHeadStem : m.tHeadBowl);
        hit |= capCapArg<S>(cap, m.headCap) < (part == 0 ? m.tStemEar : m.tBowlEar);
    }
    if (hit) return false;
    // the 16 beads of the rim and the base (:579-587)
#pragma unroll 2
    for (int j = 0; j < 2 * BEADS; ++j) {
        const S x = m.beadX[j >> 1], y = m.beadY[j >> 1], z = (j & 1) ? m.beadZBottom : m.beadZTop;
        S c[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) c[i] = R.t[i] + ((R.L[i][0] * x + R.L[i][1] * y) + R.L[i][2] * z);
#pragma unroll
        for (int k = 0; k < 3; ++k) hit |= dist2<S>(c, m.torso[k]) < m.r2TorsoBead[k];
        hit |= segPointArg<S>(m.coke, c) < m.tBeadCoke;
        hit |= segPointArg<S>(m.pepsi, c) < m.tBeadPepsi;
        hit |= dist2<S>(c, m.table) < m.r2PlaneBead;
        hit |= dist2<S>(c, m.wall) < m.r2PlaneBead;
        hit |= segPointArg<S>(Lh.arm, c) < m.tBeadArm;
        hit |= dist2<S>(c, m.headC) < m.r2HeadBead;
        hit |= segPointArg<S>(m.headCap, c) < m.tBeadEar;
    }
    return !hit;
}

template <typename S>
struct Validator {
    static constexpr int MAXD = DIM;
    static constexpr int MAXDEPTH = 10;  // below the 5 split levels: edges up to 2^15 degrees
    static constexpr bool CHECK_ENDS = false;  // nao_link assumes them (naocup.hpp:833-837)
    Model<S> model;  // ~0.5 KB (float) / 1 KB (double) of kernel parameters: read from the constant bank
    __device__ __forceinline__ int dims() const { return DIM; }
    __device__ __forceinline__ bool valid(const S* q) const { return clear<S>(model, q); }
    // nao_link_impl: d < DISCRETIZATION on nao_dist (naocup.hpp:771-783,813-815)
    __device__ __forceinline__ bool stop(const S* a, const S* b) const {
        S sum = S(0);
#pragma unroll
        for (int i = 0; i < DIM; ++i) {
            const S d = b[i] - a[i];
            sum = sum + d * d;
        }
        return fp::sqrt_(sum) < model.disc;
    }
    // bound of the recursion depth for the flat edge check (geom.cu): the ends of a node at depth k are d / 2^k apart up
    // to the rounding of k midpoints (each within an ulp of the coordinates' magnitude, 10 coordinates)
    __device__ __forceinline__ int levels(const S* a, const S* b) const {
        S sum = S(0), mag = S(1);
#pragma unroll
        for (int i = 0; i < DIM; ++i) {
            const S d = b[i] - a[i];
            sum = sum + d * d;
            mag = fmax(mag, fmax(fp::abs_(a[i]), fp::abs_(b[i])));
        }
        const S d = fp::sqrt_(sum);
        // not finite: the reference's recursion need not terminate there (a NaN in a left-arm joint compares "clear" in every
        // test of the ball while the stop test never holds); defined as an invalid edge, here and in the oracle
        if (!(d < fp::consts<S>::inf())) return -1;
        if (d < model.disc) return 0;  // the reference's own test at the root
        int L = 0;
        S x = d;
        const S t = model.disc * S(63.0 / 64.0) - S(1024) * fp::consts<S>::eps() * mag;
        if (!(t > S(0))) return 99;
        while (x >= t && L <= 24) {
            x = x * S(0.5);
            ++L;
        }
        return L;
    }
};
#endif  // __CUDACC__

}  // namespace nao
}  // namespace mptg
