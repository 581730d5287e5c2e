// builder.h -- host-side mirror of gsdf.Builder (gsdf.go:44): constructs the CSG tree table, validates shape
// parameters with the reference's rules, and computes Bounds() exactly as each Go node type does.
//
// The reference is Go; no Go toolchain exists in this image, so the host layer above the C ABI is C++ and keeps the
// reference's names and argument meaning (NewSphere, NewCylinder, Union, SmoothUnion, Translate, Extrude, ...).
// Shape errors are accumulated like Builder with FlagNoDimensionPanic (gsdf.go:100-106) and read back with Err().
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../../include/gsdf_tree.h"
#include "geom.h"

namespace gsdfhost {

using NodeId = int32_t;

class Builder {
public:
    // ---- 3D primitives (primitives.go) ----
    NodeId NewSphere(float r);                                        // :28
    NodeId NewBox(float x, float y, float z, float round);            // :65
    NodeId NewCylinder(float r, float h, float rounding);             // :107
    NodeId NewHexagonalPrism(float face2Face, float h);               // :157
    NodeId NewTriangularPrism(float triHeight, float extrudeLength);  // :198
    NodeId NewTorus(float greaterRadius, float lesserRadius);         // :216
    NodeId NewBoxFrame(float dimX, float dimY, float dimZ, float e);  // :254
    NodeId NewBoundsBoxFrame(const Box3 &bb);                         // :12
    // ---- 3D operations (operations.go) ----
    NodeId Union(const std::vector<NodeId> &shaders);                 // :35 (flattens nested unions)
    NodeId Difference(NodeId a, NodeId b);                            // :117
    NodeId Intersection(NodeId a, NodeId b);                          // :160
    NodeId Xor(NodeId a, NodeId b);                                   // :205
    NodeId Scale(NodeId s, float scaleFactor);                        // :248
    NodeId Symmetry(NodeId s, bool mx, bool my, bool mz);             // :285
    NodeId Transform(NodeId s, const Mat4 &m);                        // :340
    NodeId Rotate(NodeId s, float radians, Vec3 axis);                // :394
    NodeId Translate(NodeId s, float dx, float dy, float dz);         // :403
    NodeId Offset(NodeId s, float sdfAdd);                            // :446
    NodeId Array(NodeId s, float sx, float sy, float sz, int nx, int ny, int nz);  // :488
    NodeId SmoothUnion(float k, NodeId a, NodeId b);                  // :563
    NodeId SmoothDifference(float k, NodeId a, NodeId b);             // :611
    NodeId SmoothIntersect(float k, NodeId a, NodeId b);              // :643
    NodeId Elongate(NodeId s, float dx, float dy, float dz);          // :679
    NodeId Shell(NodeId s, float thickness);                          // :723
    NodeId CircularArray(NodeId s, int numInstances, int circleDiv);  // :764
    NodeId Twist(NodeId s, float k);                                  // :835
    NodeId OverloadShader3DBounds(NodeId s, const Box3 &bb);          // glbuild/glbuild.go:1080
    NodeId OverloadShader2DBounds(NodeId s, const Box2 &bb);          // glbuild/glbuild.go:1105
    // ---- 2D -> 3D (operations2d.go) ----
    NodeId Extrude(NodeId s2, float h);                               // :104
    NodeId Revolve(NodeId s2, float axisOffset);                      // :149
    // forge/threads/threads.go:76-96 (raw screw node; Threader front-ends live in threads.h)
    NodeId NewScrew(NodeId thread2d, float pitch, float lead, float length, float taper);
    // ---- 2D primitives (primitives2d.go) ----
    NodeId NewLine2D(float x0, float y0, float x1, float y1, float width);  // :14
    NodeId NewLines2D(const std::vector<Vec2> &segPairs, float width);      // :62 (2 points per segment)
    NodeId NewArc(float radius, float arcAngle, float thick);               // :169
    NodeId NewCircle(float r);                                              // :227
    NodeId NewEquilateralTriangle(float h);                                 // :265
    NodeId NewRectangle(float x, float y);                                  // :307
    NodeId NewHexagon(float side);                                          // :348
    NodeId NewOctagon(float c);                                             // :385
    NodeId NewPolygon(std::vector<Vec2> vertices);                          // :458 (+ validatePolygon :471)
    NodeId NewDiamond2D(float w, float h);                                  // :560
    NodeId NewRoundedX(float width, float thick);                           // :602
    NodeId NewEllipse(float a, float b);                                    // :421
    NodeId NewQuadraticBezier2D(Vec2 a, Vec2 b, Vec2 c, float thick);       // :644
    // ---- 2D operations (operations2d.go) ----
    NodeId Union2D(const std::vector<NodeId> &shaders);               // :18
    NodeId Difference2D(NodeId a, NodeId b);                          // :201
    NodeId Intersection2D(NodeId a, NodeId b);                        // :245
    NodeId Xor2D(NodeId a, NodeId b);                                 // :289
    NodeId Array2D(NodeId s, float sx, float sy, int nx, int ny);     // :332
    NodeId Offset2D(NodeId s, float sdfAdd);                          // :410
    NodeId Translate2D(NodeId s, float dx, float dy);                 // :456
    NodeId Rotate2D(NodeId s, float theta);                           // :494
    NodeId Symmetry2D(NodeId s, bool mx, bool my);                    // :555
    NodeId Annulus(NodeId s, float sub);                              // :604
    NodeId CircularArray2D(NodeId s, int numInstances, int circleDiv);// :655
    NodeId Scale2D(NodeId s, float scale);                            // :719
    NodeId TranslateMulti2D(NodeId s, const std::vector<Vec2> &disp); // :756
    NodeId Elongate2D(NodeId s, float dx, float dy);                  // :826

    // ---- queries ----
    bool is3D(NodeId id) const;
    bool is2D(NodeId id) const;
    bool valid(NodeId id) const { return id >= 0 && id < (NodeId)nodes_.size(); }
    Box3 Bounds3(NodeId id) const;  // each node type's Bounds() (cited in builder.cpp)
    Box2 Bounds2(NodeId id) const;
    std::string Err() const;        // gsdf.go:88 -- accumulated shape errors, "" if none
    void ClearErrors() { errs_.clear(); }

    const std::vector<gsdf_tree_node> &nodes() const { return nodes_; }
    const std::vector<int32_t> &children() const { return children_; }
    const std::vector<float> &aux() const { return aux_; }
    const gsdf_tree_node &node(NodeId id) const { return nodes_[id]; }
    NodeId child(const gsdf_tree_node &n, int k) const { return children_[n.child_off + k]; }

private:
    NodeId push(int kind, std::initializer_list<float> f, const std::vector<NodeId> &ch, std::initializer_list<int> ip = {});
    void shapeErrorf(const std::string &msg) { errs_.push_back(msg); }
    bool need3(NodeId s, const char *who);
    bool need2(NodeId s, const char *who);

    std::vector<gsdf_tree_node> nodes_;
    std::vector<int32_t> children_;
    std::vector<float> aux_;
    std::vector<std::string> errs_;
};

}  // namespace gsdfhost
