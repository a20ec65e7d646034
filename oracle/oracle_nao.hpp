// oracle_nao.hpp -- CPU restatement of the reference's Nao-cup scenario.
// TEST INFRASTRUCTURE ONLY (see oracle.hpp header): the product never includes this file.
//
// Follows demo/nao_cup/src/naocup.hpp, collide.hpp, linear.hpp and demo/nao_cup_planning.cpp:146-152:
//   valid(q) = nao_clear  = cup_is_up && !in_collision         (naocup.hpp:795-806, :708, :728)
//   link(a,b) = nao_link  = midpoint bisection down to 1 degree, endpoints not checked (naocup.hpp:809-840)
// The forward kinematics are kept GENERAL here -- 3x3 linear part + translation, every rotation built by the
// angle-axis formula and applied by a full matrix product, every translation by a full matrix-vector product, in the
// operation structure of Eigen's Transform (T*AngleAxis: linear *= R; T*Translation: translation += linear*v) with
// sums running left to right -- while the CUDA kernel (mpt_b200/csrc/nao.cuh) uses closed forms with the exact zeros
// and ones removed.  Parity of the two is therefore a real check of both.
// Pinned against the reference's own code compiled here (oracle/ref_nao.cpp -> oracle/_ref/libref_nao.so, golden
// vectors in tests/golden/nao_golden.npz); the reference calls libm sin/cos, this file mptg_fpmath.h.
#pragma once

#include <cmath>
#include <cstdint>
#include <limits>

#include "../include/mptg/mptg_fpmath.h"

namespace oracle {

template <typename S>
struct NaoCup {
    static constexpr int DIM = 10;  // naocup.hpp:54; joints R/L shoulder pitch, shoulder roll, elbow yaw, elbow roll, wrist yaw (:178-190)
    struct Xf {                     // Transform<S,3,Isometry> (linear.hpp:44)
        S L[3][3], t[3];
    };
    struct Obj {  // collision_object (naocup.hpp:88-93)
        bool capsule;
        const Xf* xf;
        S radius, length;
    };

    // ---- constants, written with the reference's expression types (naocup.hpp:55-80,222-251)
    static S inch() { return S(0.0254); }
    static S pi() { return S(3.14159265358979323846); }
    static S rhandUpGoal() { return S(0.90630778703665); }
    static S centerTorsoRadius() { return S(66.7 / 1000.0); }
    static constexpr int BEADS = 8;
    static S beadRadius() { return S(2.0 / 4.0) * inch(); }
    static S cupDiameter() { return S(2.5) * inch(); }
    static S cupHeight() { return (S(4.0) + S(3.0) / S(8.0)) * inch(); }
    static S gripHeight() { return S(1.0) * inch(); }
    static S baseToBowl() { return (S(1.0) + S(5.0) / S(8.0)) * inch(); }
    static S gripDiameter() { return S(5.0) / S(8.0) * inch(); }
    static S gripCapsuleHeight() { return baseToBowl() - gripDiameter() * S(2.0); }
    static S bowlHeight() { return cupHeight() - baseToBowl(); }
    static S ballRadius() { return S(0.015); }
    static S planarRadius() { return S(25.0); }
    static S tableZ() { return S(0.09); }
    static S discretization() { return S(1.0) * pi() / S(180.0); }
    static S neckZ() { return S(126.50 / 1000.0); }
    static S shoulderY() { return S(98.00 / 1000.0); }
    static S upperArm() { return S(90.00 / 1000.0); }
    static S lowerArm() { return S(50.55 / 1000.0); }
    static S shoulderZ() { return S(100.00 / 1000.0); }
    static S handX() { return S(58.00 / 1000.0); }
    static S hipZ() { return S(85.00 / 1000.0); }
    static S handZ() { return S(15.90 / 1000.0); }
    static S headRadius() { return S(115.0 / 2.0 / 1000.0); }
    static S earRadius() { return S(90.0 / 2.0 / 1000.0); }
    static S headWidth() { return S(133.0 / 1000.0); }
    static S armRadius() { return S(66.7 / 2.0 / 1000.0); }
    static S handRadius() { return S(20.0 / 1000.0); }
    static S handWidth() { return S(50.0 / 1000.0); }

    // ---- linear.hpp:74-95 on the Transform structure stated above
    static Xf identity() {
        Xf r;
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) r.L[i][j] = S(i == j ? 1 : 0);
            r.t[i] = S(0);
        }
        return r;
    }
    static Xf rotate(const Xf& a, S angle, S x, S y, S z) {  // m4_rotate: *t * AngleAxis(a, (x,y,z))
        S sn, cs;
        mptg::fp::sincos_(angle, &sn, &cs);
        const S ax[3] = {x, y, z};
        const S sx = sn * ax[0], sy = sn * ax[1], sz = sn * ax[2];
        const S cx = (S(1) - cs) * ax[0], cy = (S(1) - cs) * ax[1], cz = (S(1) - cs) * ax[2];
        S R[3][3], tmp;
        tmp = cx * ax[1], R[0][1] = tmp - sz, R[1][0] = tmp + sz;
        tmp = cx * ax[2], R[0][2] = tmp + sy, R[2][0] = tmp - sy;
        tmp = cy * ax[2], R[1][2] = tmp - sx, R[2][1] = tmp + sx;
        R[0][0] = cx * ax[0] + cs, R[1][1] = cy * ax[1] + cs, R[2][2] = cz * ax[2] + cs;
        Xf o;
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) o.L[i][j] = (a.L[i][0] * R[0][j] + a.L[i][1] * R[1][j]) + a.L[i][2] * R[2][j];
            o.t[i] = a.t[i];
        }
        return o;
    }
    static Xf translate(const Xf& a, S x, S y, S z) {  // m4_translate: *t * Translation(x,y,z)
        Xf o = a;
        for (int i = 0; i < 3; ++i) o.t[i] = a.t[i] + ((a.L[i][0] * x + a.L[i][1] * y) + a.L[i][2] * z);
        return o;
    }
    static void apply(const Xf& m, S x, S y, S z, S out[3]) {  // m4_transform_i3 / the Vec3 overload of m4_transform_i
        for (int i = 0; i < 3; ++i) out[i] = ((m.L[i][0] * x + m.L[i][1] * y) + m.L[i][2] * z) + m.t[i];
    }
    static void apply4(const Xf& m, S x, S y, S z, S w, S out[3]) {  // m4_transform_i, Vec4 overload (first three rows)
        for (int i = 0; i < 3; ++i) out[i] = ((m.L[i][0] * x + m.L[i][1] * y) + m.L[i][2] * z) + m.t[i] * w;
    }
    static S norm3(const S v[3]) { return mptg::fp::sqrt_((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]); }
    static S dot3(const S a[3], const S b[3]) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
    static S distSegmentPoint(const S s0[3], const S s1[3], const S pt[3]) {  // linear.hpp:125-150
        S v[3], w[3];
        for (int i = 0; i < 3; ++i) v[i] = s1[i] - s0[i], w[i] = pt[i] - s0[i];
        const S c1 = dot3(w, v);
        if (c1 <= S(0)) return norm3(w);
        const S c2 = dot3(v, v);
        if (c2 <= c1) {
            S e[3];
            for (int i = 0; i < 3; ++i) e[i] = pt[i] - s1[i];
            return norm3(e);
        }
        const S b = c1 / c2;
        S e[3];
        for (int i = 0; i < 3; ++i) {
            const S p = s0[i] + v[i] * b;
            e[i] = p - pt[i];
        }
        return norm3(e);
    }

    // ---- collide.hpp:45-115.  *margin (optional) is lowered to |distance - reach| of the closest call.
    static void note(double* margin, double d, double r) {
        if (margin) *margin = std::fmin(*margin, std::fabs(d - r));
    }
    static bool sphereCapsule(const Xf& st, S sr, const Xf& ct, S len, S cr, double* margin) {
        S c[3], p0[3], p1[3];
        apply(st, S(0), S(0), S(0), c);
        apply(ct, S(0), S(0), S(0), p0);
        apply(ct, S(0), S(0), len, p1);
        const S dist = distSegmentPoint(p0, p1, c);
        note(margin, (double)dist, (double)(sr + cr));
        return dist < (sr + cr);
    }
    static bool sphereSphere(const Xf& at, S ar, const Xf& bt, S br, double* margin) {
        S a[3], b[3], d[3];
        apply4(at, S(0), S(0), S(0), S(1), a);
        apply4(bt, S(0), S(0), S(0), S(1), b);
        for (int i = 0; i < 3; ++i) d[i] = a[i] - b[i];
        const S d2 = (d[0] * d[0] + d[1] * d[1]) + d[2] * d[2];
        S r2 = ar + br;
        r2 = r2 * r2;
        note(margin, std::sqrt((double)d2), std::sqrt((double)r2));
        return d2 < r2;
    }
    static bool capsuleCapsule(const Xf& at, S al, S ar, const Xf& bt, S bl, S br, double* margin) {
        S a0[3], a1[3], b0[3], b1[3];
        apply(at, S(0), S(0), S(0), a0);
        apply(at, S(0), S(0), al, a1);
        apply(bt, S(0), S(0), S(0), b0);
        apply(bt, S(0), S(0), bl, b1);
        const S da0 = distSegmentPoint(a0, a1, b0), da1 = distSegmentPoint(a0, a1, b1);
        const S db0 = distSegmentPoint(b0, b1, a0), db1 = distSegmentPoint(b0, b1, a1);
        const S dist = std::fmin(std::fmin(da0, da1), std::fmin(db0, db1));
        note(margin, (double)dist, (double)(ar + br));
        return dist < (ar + br);
    }
    static bool collideObjects(const Obj& a, const Obj& b, double* margin) {  // naocup.hpp:304-330
        if (!a.capsule && !b.capsule) return sphereSphere(*a.xf, a.radius, *b.xf, b.radius, margin);
        if (!a.capsule && b.capsule) return sphereCapsule(*a.xf, a.radius, *b.xf, b.length, b.radius, margin);
        if (a.capsule && !b.capsule) return sphereCapsule(*b.xf, b.radius, *a.xf, a.length, a.radius, margin);
        return capsuleCapsule(*a.xf, a.length, a.radius, *b.xf, b.length, b.radius, margin);
    }
    static bool collideLists(const Obj* a, int n, const Obj* b, int m, double* margin, uint64_t* pairs) {  // :333-349 (no early out)
        bool hit = false;
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < m; ++j) {
                if (collideObjects(a[i], b[j], margin)) hit = true;
                if (pairs) ++*pairs;
            }
        return hit;
    }

    struct Arm {
        Xf lowerCapsule, toHand;
    };
    static Arm computeArm(const Xf& base, const S* c) {  // compute_arm, naocup.hpp:375-406 (the unused upper capsule is left out)
        const Xf shoulderPitch = rotate(base, c[0], S(0), S(1), S(0));
        const Xf shoulderRoll = rotate(shoulderPitch, c[1], S(0), S(0), S(1));
        const Xf toElbow = translate(shoulderRoll, upperArm(), S(0), S(0));
        const Xf elbowYaw = rotate(toElbow, c[2], S(1), S(0), S(0));
        const Xf elbowRoll = rotate(elbowYaw, c[3], S(0), S(0), S(1));
        Arm arm;
        arm.lowerCapsule = rotate(elbowRoll, pi() / S(2.0), S(0), S(1), S(0));
        const Xf wristYaw = rotate(elbowRoll, c[4], S(1), S(0), S(0));
        const Xf toHand = translate(wristYaw, lowerArm() + handX() + S(0.01), S(0), -handZ() - S(0.01));
        const Xf rotateHand = rotate(toHand, -pi() / S(2.0), S(1), S(0), S(0));
        arm.toHand = translate(rotateHand, S(0), S(0), (handWidth() - handRadius() * S(2)) / S(2));
        return arm;
    }

    // compute() + check_collisions() (naocup.hpp:565-730), the part nao_clear reads.
    // margin (optional): smallest |distance - reach| over the pair tests and |cup axis z - threshold|.
    bool valid(const S* q, bool* collisionOut = nullptr, double* margin = nullptr, uint64_t* pairs = nullptr) const {
        if (margin) *margin = std::numeric_limits<double>::infinity();
        const Xf robot = identity();
        // head: yaw and pitch are fields of the value-initialised world that nothing sets (naocup.hpp:118-119, :866): zero
        const Xf headBase = translate(robot, S(0), S(0), neckZ());
        const Xf headYaw = rotate(headBase, S(0), S(0), S(0), S(1));
        const Xf headPitch = rotate(headYaw, S(0), S(0), S(1), S(0));
        const Xf headCenter = translate(headPitch, S(0), S(0), headRadius());
        const Xf headCapsuleRotate = rotate(headCenter, pi() / S(2.0), S(1), S(0), S(0));
        const Xf headCapsule = translate(headCapsuleRotate, S(0), S(0), -(headWidth() - earRadius() * S(2)) / S(2));
        const Xf upperTorso = translate(robot, S(0), S(0), neckZ() - S(66.7) / S(1000.0));
        const Xf lowerTorso = translate(robot, S(0), S(0), -hipZ() / S(2.0));
        const Arm right = computeArm(translate(robot, S(0), -shoulderY(), shoulderZ()), q);
        const Arm left = computeArm(translate(robot, S(0), shoulderY(), shoulderZ()), q + 5);

        // check_collisions, :565-586
        const Xf coke = translate(robot, S(0.12), S(0.08), -tableZ());
        const Xf pepsi = translate(robot, S(0.12) + S(2.5) * inch(), S(-0.12), -tableZ());
        const Xf table = translate(robot, S(0.0), S(0.0), -planarRadius() - tableZ());
        const Xf backwall = translate(robot, -planarRadius() - centerTorsoRadius(), S(0.0), S(0.0));
        const Xf ballXf = translate(left.toHand, S(0.0), ballRadius() / S(2.0) + S(0.01), S(0.0));
        const Xf cupStem = translate(right.toHand, S(0.0), S(0.0), gripHeight() / S(2.0));
        const Xf cupBowl = translate(right.toHand, S(0.0), S(0.0), gripHeight() / S(2.0) + cupDiameter() / S(2.0));
        Xf bead[BEADS * 2];
        for (int i = 0; i < BEADS; ++i) {
            S x, y;
            beadOffset(i, &x, &y);
            bead[i * 2] = translate(right.toHand, x, y, bowlHeight() + gripHeight() / S(2.0) - beadRadius());
            bead[i * 2 + 1] = translate(right.toHand, x, y, gripHeight() / S(2.0) - baseToBowl() + beadRadius());
        }
        // init_collisions, :426-553
        const Obj rightArm[1] = {{true, &right.lowerCapsule, armRadius(), lowerArm() + handX() - armRadius()}};
        const Obj leftArm[1] = {{true, &left.lowerCapsule, armRadius(), lowerArm() + handX() - armRadius()}};
        const Obj torso[3] = {{false, &robot, S(55.6 / 1000.0), S(0)},
                              {false, &upperTorso, centerTorsoRadius(), S(0)},
                              {false, &lowerTorso, hipZ() / S(2.0), S(0)}};
        const Obj head[2] = {{false, &headCenter, headRadius(), S(0)}, {true, &headCapsule, earRadius(), headWidth() - earRadius() / S(2.0)}};
        const Obj obstacles[4] = {{true, &coke, S(2.5) / S(2.0) * inch(), (S(6.75) - S(2.5) / S(2.0)) * inch()},
                                  {true, &pepsi, S(3.0) / S(2.0) * inch(), (S(8.5) - S(3.0) / S(2.0)) * inch()},
                                  {false, &table, planarRadius(), S(0)},
                                  {false, &backwall, planarRadius(), S(0)}};
        const Obj ball[1] = {{false, &ballXf, ballRadius(), S(0)}};
        Obj cup[BEADS * 2 + 2];
        cup[0] = {true, &cupStem, gripDiameter() / S(2.0), -gripCapsuleHeight()};
        cup[1] = {true, &cupBowl, cupDiameter() / S(2.0), bowlHeight() - cupDiameter()};
        for (int i = 0; i < BEADS * 2; ++i) cup[i + 2] = {false, &bead[i], beadRadius(), S(0)};
        constexpr int NC = BEADS * 2 + 2;
        // :592-613, every list evaluated
        int hits = 0;
        hits += collideLists(torso, 3, rightArm, 1, margin, pairs) ? 1 : 0;
        hits += collideLists(torso, 3, leftArm, 1, margin, pairs) ? 1 : 0;
        hits += collideLists(leftArm, 1, rightArm, 1, margin, pairs) ? 1 : 0;
        hits += collideLists(head, 2, rightArm, 1, margin, pairs) ? 1 : 0;
        hits += collideLists(head, 2, leftArm, 1, margin, pairs) ? 1 : 0;
        hits += collideLists(rightArm, 1, obstacles, 4, margin, pairs) ? 1 : 0;
        hits += collideLists(leftArm, 1, obstacles, 4, margin, pairs) ? 1 : 0;
        hits += collideLists(cup, NC, torso, 3, margin, pairs) ? 1 : 0;
        hits += collideLists(cup, NC, obstacles, 4, margin, pairs) ? 1 : 0;
        hits += collideLists(cup, NC, leftArm, 1, margin, pairs) ? 1 : 0;
        hits += collideLists(cup, NC, head, 2, margin, pairs) ? 1 : 0;
        hits += collideLists(rightArm, 1, ball, 1, margin, pairs) ? 1 : 0;
        hits += collideLists(head, 2, ball, 1, margin, pairs) ? 1 : 0;
        hits += collideLists(torso, 3, ball, 1, margin, pairs) ? 1 : 0;
        hits += collideLists(obstacles, 4, ball, 1, margin, pairs) ? 1 : 0;
        const bool inCollision = hits != 0;
        if (collisionOut) *collisionOut = inCollision;

        // cup_is_up, :673-689
        S rPos[3], bowlCenter[3], cupUp[3];
        for (int i = 0; i < 3; ++i) rPos[i] = right.toHand.t[i];
        apply(right.toHand, S(0), S(0), S(bowlHeight() + gripHeight() / S(2.0)), bowlCenter);
        for (int i = 0; i < 3; ++i) cupUp[i] = bowlCenter[i] - rPos[i];
        const S n2 = (cupUp[0] * cupUp[0] + cupUp[1] * cupUp[1]) + cupUp[2] * cupUp[2];
        const S upZ = n2 > S(0) ? cupUp[2] / mptg::fp::sqrt_(n2) : cupUp[2];  // Eigen normalized()
        if (margin) *margin = std::fmin(*margin, std::fabs((double)upZ - (double)rhandUpGoal()));
        const bool cupIsUp = upZ > rhandUpGoal();
        return cupIsUp && !inCollision;  // is_clear, :728
    }

    // bead centres on the rim (naocup.hpp:540-543): the angle is formed in S, cos/sin are the double-precision
    // functions (the unqualified calls resolve to ::cos(double) whatever S is), the product is formed in double
    static void beadOffset(int i, S* x, S* y) {
        const S a = pi() * S(2.0) * (S)i / (S)BEADS;
        double sn, cs;
        mptg::fp::sincos_((double)a, &sn, &cs);
        const S r = cupDiameter() / S(2.0) - beadRadius();
        *x = (S)(cs * (double)r);
        *y = (S)(sn * (double)r);
    }

    static S dist(const S* a, const S* b) {  // nao_dist, naocup.hpp:771-783
        S sum = 0;
        for (int i = 0; i < DIM; ++i) {
            const S d = b[i] - a[i];
            sum = sum + d * d;
        }
        return mptg::fp::sqrt_(sum);
    }
    bool linkImpl(const S* a, const S* b, uint64_t* states) const {  // nao_link_impl, :809-825
        const S d = dist(a, b);
        // A NaN or infinite joint value: the reference's recursion need not terminate (a NaN in a left-arm joint makes
        // every comparison of the ball's tests false, i.e. "clear", while the stop test never holds).  Defined here and
        // in the kernels: an edge whose length is not finite is invalid.
        if (!(d < std::numeric_limits<S>::infinity())) return false;
        if (d < discretization()) return true;
        S m[DIM];
        for (int i = 0; i < DIM; ++i) m[i] = (a[i] + b[i]) / S(2.0);
        if (states) ++*states;
        return valid(m) && linkImpl(a, m, states) && linkImpl(m, b, states);
    }
    bool link(const S* a, const S* b, uint64_t* states = nullptr) const { return linkImpl(a, b, states); }  // :828-840: ends not checked
    // every midpoint of the recursion regardless of the outcome: smallest margin along the edge (test support)
    void linkMargin(const S* a, const S* b, double* margin) const {
        const S d = dist(a, b);
        if (!(d < std::numeric_limits<S>::infinity()) || d < discretization()) return;
        S m[DIM];
        for (int i = 0; i < DIM; ++i) m[i] = (a[i] + b[i]) / S(2.0);
        double mm;
        valid(m, nullptr, &mm);
        *margin = std::fmin(*margin, mm);
        linkMargin(a, m, margin);
        linkMargin(m, b, margin);
    }
};

}  // namespace oracle
