"""B200-native hot path of ebwt2InDel: eBWT suffix-tree traversal -> LCP bitvectors -> .snp calls.

Python here is plumbing only (ctypes bindings to the C-ABI library built from csrc/, synthetic
input tooling for tests and bench).  The product is include/e2i.h + libe2i.so + bin/ebwt2InDel.
"""
__version__ = "0.1.0"
