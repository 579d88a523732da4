// Umbrella header (reference: huffman/huffman.hpp:3-8).
#pragma once

#include "huffman/src/bit.hpp"
#include "huffman/src/bit_span.hpp"
#include "huffman/src/code.hpp"
#include "huffman/src/decode.hpp"
#include "huffman/src/symbol_span.hpp"
#include "huffman/src/table.hpp"
#include "huffman/src/utility.hpp"
