/* Empty stand-in for the legacy <THC/THC.h> header, which PyTorch >= 1.11 no longer ships and the
 * reference's tree_filter sources still include (bfs.cu:12, refine.cu:12, mst.cu:9) without using
 * anything from it.  It lets oracle/Makefile compile those sources UNMODIFIED, from where they lie. */
#pragma once
