// TEST INFRASTRUCTURE ONLY.  PothosCore is absent from this image; the reference's
// filter/FIRFilter.cpp is compiled against the repo's own restatement of the Pothos API subset it
// touches, the same header the product's block layer builds against.
#pragma once
#include "../../../pothoscomms_b200/blocks/shim/Pothos/Framework.hpp"
