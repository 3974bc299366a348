/* A plain-C caller of the drop-in boundary (what bindgen's `extern "C"` block resolves to in the reference,
 * graph/src/graph/graphblas/mod.rs).  Host-only part: element ops, pending work, row iterator, container round trip --
 * runs anywhere.  Device part (argument "gpu"): the batched traversal F <- F*A*A over GxB_ANY_PAIR_BOOL with the result
 * walked by the row iterator, exactly the calls CondTraverse makes (cond_traverse.rs:600-653).
 *   gcc -std=c11 -I include examples/c_abi_demo.c -L falkordb_b200 -lb200grb -Wl,-rpath,$PWD/falkordb_b200 -o demo */
#include "b200grb.h"
#include <stdio.h>
#include <string.h>

#define OK(x) do { GrB_Info i_ = (x); if (i_ != GrB_SUCCESS) { fprintf(stderr, "%s -> %d (%s)\n", #x, (int)i_, B200_last_error()); return 1; } } while (0)

static int dump(GrB_Matrix A, const char *name) {
    GxB_Iterator it;
    OK(GxB_Iterator_new(&it));
    OK(GxB_rowIterator_attach(it, A, NULL));
    GrB_Info info = GxB_rowIterator_seekRow(it, 0);
    printf("%s:", name);
    while (info != GxB_EXHAUSTED) {
        if (info == GrB_SUCCESS) {
            do {
                printf(" (%llu,%llu)", (unsigned long long)GxB_rowIterator_getRowIndex(it), (unsigned long long)GxB_rowIterator_getColIndex(it));
            } while (GxB_rowIterator_nextCol(it) == GrB_SUCCESS);
        }
        info = GxB_rowIterator_nextRow(it);
    }
    printf("\n");
    return (int)GxB_Iterator_free(&it);
}

int main(int argc, char **argv) {
    OK(GxB_init(GrB_NONBLOCKING, NULL, NULL, NULL, NULL));
    /* the README MotoGP shape: riders 0..2 -> teams 3..5 (README.md:85-110) */
    GrB_Matrix rides;
    OK(GrB_Matrix_new(&rides, GrB_BOOL, 6, 6));
    OK(GrB_Matrix_setElement_BOOL(rides, true, 0, 3));
    OK(GrB_Matrix_setElement_BOOL(rides, true, 1, 4));
    OK(GrB_Matrix_setElement_BOOL(rides, true, 2, 5));
    OK(GrB_Matrix_setElement_BOOL(rides, true, 2, 5));            /* a duplicate write is one entry */
    OK(GrB_Matrix_removeElement(rides, 1, 4));
    OK(GrB_Matrix_wait(rides, GrB_MATERIALIZE));
    GrB_Index nv;
    OK(GrB_Matrix_nvals(&nv, rides));
    printf("nvals %llu\n", (unsigned long long)nv);
    if (dump(rides, "rides")) return 1;

    /* RDB save / restore through the container (matrix.rs:428-546) */
    GxB_Container c;
    OK(GxB_Container_new(&c));
    OK(GxB_unload_Matrix_into_Container(rides, c, NULL));
    printf("container %llux%llu nvals %llu format %d iso %d\n", (unsigned long long)c->nrows, (unsigned long long)c->ncols,
           (unsigned long long)c->nvals, c->format, (int)c->iso);
    GrB_Matrix back;
    OK(GrB_Matrix_new(&back, GrB_BOOL, 0, 0));
    OK(GxB_load_Matrix_from_Container(back, c, NULL));
    OK(GxB_Container_free(&c));
    if (dump(back, "restored")) return 1;

    if (argc > 1 && !strcmp(argv[1], "gpu")) {
        /* F(i, src_i) = 1 ; F <- F * rides * rides' : riders that share a team with rider i */
        GrB_Matrix F, T;
        GrB_Index rows[2] = {0, 1}, cols[2] = {0, 2};
        GrB_Scalar one;
        OK(GrB_Scalar_new(&one, GrB_BOOL));
        OK(GrB_Scalar_setElement_BOOL(one, true));
        OK(GrB_Matrix_new(&F, GrB_BOOL, 2, 6));
        OK(GxB_Matrix_build_Scalar(F, rows, cols, one, 2));
        OK(GrB_Matrix_new(&T, GrB_BOOL, 6, 6));
        OK(GrB_transpose(T, NULL, NULL, back, NULL));
        OK(GrB_mxm(F, NULL, NULL, GxB_ANY_PAIR_BOOL, F, back, NULL));
        OK(GrB_mxm(F, NULL, NULL, GxB_ANY_PAIR_BOOL, F, T, NULL));
        OK(GrB_Matrix_wait(F, GrB_MATERIALIZE));
        if (dump(F, "teammates")) return 1;
        GrB_Matrix_free(&F); GrB_Matrix_free(&T); GrB_Scalar_free(&one);
    }
    GrB_Matrix_free(&rides); GrB_Matrix_free(&back);
    return (int)GrB_finalize();
}
