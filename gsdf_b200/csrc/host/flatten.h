// flatten.h -- tree table -> packed node program (include/gsdf_program.h). See flatten.cpp.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../../include/gsdf_program.h"
#include "builder.h"

namespace gsdfhost {

struct Program {
    std::vector<uint32_t> chunks;  // 4 words per 16-byte chunk, END included
    std::vector<float> aux;        // side buffer (polygon edge records, line segments)
    int dim = 3;
    int dstack = 1, pstack = 0, ninstr = 0;
    std::vector<uint8_t> blob() const;  // gsdf_program_header + chunks, ready for gsdf_program_create
};

// Replaces glbuild.Programmer.WriteComputeSDF3/2 (glbuild/glbuild.go:175,218) for the CUDA backend.
bool Flatten(const Builder &b, NodeId root, Program &out, std::string &err);

}  // namespace gsdfhost
