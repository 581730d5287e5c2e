// threads.h -- host-side mirror of the reference's forge/threads package (screw node front-ends and the
// Bolt / Nut / HexHead / Knurl builders) plus the example scenes BASELINE.json names. These are scene
// CONSTRUCTION code (they run once on the host); the hot path consumes the tree they produce.
#pragma once
#include <string>

#include "builder.h"

namespace gsdfhost {
namespace threads {

// forge/threads/threads.go:33-51
struct Parameters {
    std::string Name;
    float Radius = 0, Pitch = 0;
    int Starts = 1;
    float Taper = 0, HexF2F = 0;
    float HexRadius() const;
    float HexHeight() const;
};

enum class Kind { ISO, NPT, Knurl, UTS, Acme, ANSIButtress, PlasticButtress };  // uts.go, acme.go, ansibuttress.go, plasticbuttress.go
// One value type covers the reference's Threader implementations used by the benchmark scenes.
struct Threader {
    Kind kind = Kind::ISO;
    // ISO (iso.go:20-29) / NPT (npt.go:11-19)
    float D = 0, P = 0;
    bool Ext = true;
    float TPI = 0, F2F = 0;
    // KnurlParams (knurl.go:18-25)
    float KLength = 0, KRadius = 0, KPitch = 0, KHeight = 0, KTheta = 0;
    int kstarts = 0;

    static Threader ISO(float D, float P, bool ext);
    static Threader Basic(Kind kind, float D, float P);       // Acme / ANSIButtress / PlasticButtress {D, P}
    static Threader UTS(float D, float TPI, bool ext);        // uts.go:8-15
    static bool NPTFromNominal(float nominal, Threader &out);  // npt.go:63-74
    Parameters ThreadParams() const;                            // iso.go:33, npt.go:23, knurl.go:45
    NodeId Thread(Builder &bld, std::string &err) const;        // iso.go:37, npt.go:34, knurl.go:28
};

enum NutStyle { NutCircular = 1, NutHex, NutKnurl };  // nut.go:12-17

NodeId Screw(Builder &bld, float length, const Threader &t, std::string &err);                                // threads.go:76
NodeId HexHead(Builder &bld, float radius, float height, bool roundNeg, bool roundPos, std::string &err);     // hexhead.go:15
NodeId Knurl(Builder &bld, Threader k, std::string &err);                                                     // knurl.go:51
NodeId KnurledHead(Builder &bld, float radius, float height, float pitch, std::string &err);                  // knurl.go:84
NodeId Nut(Builder &bld, const Threader &t, NutStyle style, float tolerance, std::string &err);               // nut.go:41
NodeId Bolt(Builder &bld, const Threader &t, NutStyle style, float tolerance, float totalLength, float shankLength,
            std::string &err);                                                                                 // bolt.go:21
float metricf2f(float radius);                                                                                 // threads.go:229

}  // namespace threads

namespace scenes {
NodeId NptFlange(Builder &bld, std::string &err);                 // examples/npt-flange/flange.go:23-59
NodeId Bolt(Builder &bld, std::string &err);                      // examples/bolt/main.go:26-41
NodeId KnurledCylinder(Builder &bld, float diameter, std::string &err);  // examples/knurled-cylinder/knurled-cyl.go:57-107
NodeId FibonacciShowerhead(Builder &bld, std::string &err);       // examples/fibonacci-showerhead/showerhead.go:31-92
}  // namespace scenes

}  // namespace gsdfhost
