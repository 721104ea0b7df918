/*
 * oracle/mmqr_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's MMQR sliding-window Householder QR
 * (brian-kelley/CUDA-QR, qr.c) with the window geometry (PR, PC) as run-time
 * arguments.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
 * leg may load this file's shared object; the shipped library never does.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement
 *   (a) against the reference's only known-answer run (6x4, srand(12),
 *       qr.c:461-523; values in tests/golden/demo_6x4.json), and
 *   (b) BIT FOR BIT against the unmodified reference compiled into
 *       oracle/_ref/ (oracle/build_ref.sh) on seeded inputs, for mmqr and
 *       explicitQR, at three (PR, PC) settings.
 *
 * Every floating-point operation below is performed in the same order and
 * precision as the reference so (b) can hold exactly; what is dropped is only
 * work whose result is an exact zero (multiplications by the structural zeros
 * of Y, W and H) plus all printing and per-window malloc traffic.
 *
 * Build: gcc -O2 -std=c99 -ffp-contract=off -shared -fPIC (oracle/Makefile).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

/* qr.c:45 */
static int ceil_div(int a, int b) { return a / b + (a % b != 0); }

/* qr.c:47-53 -- number of column blocks and of overlapping row windows. */
ORACLE_API void oracle_getPanelDims(int m, int n, int PR, int PC,
                                    int *rowPanels, int *colPanels)
{
    *colPanels = ceil_div(n, PC);
    *rowPanels = 1;
    if (m > PR)
        *rowPanels += ceil_div(m - PR, PR - PC);
}

/* qr.c:109-141 (also 361-396): local row range [vs, ve) of reflector `col`
 * inside the window whose top-left corner is (pr, pc). */
static void reflector_span(int m, int PR, int PC, int pr, int pc, int col,
                           int *vs, int *ve)
{
    int bottom = (pr == m - PR);
    int top = (pr <= pc);
    *vs = top ? pc - pr + col : col;
    *ve = bottom ? PR : PR - PC + col + 1;
}

/* A window's working set (column-major, leading dimension PR). */
typedef struct {
    float *panel; /* PR x PC  : the window of A              (qr.c:77)  */
    float *W;     /* PR x PC  : W of  Q_window^T = I + Y W^T  (qr.c:98)  */
    float *Y;     /* PR x PC  : explicit reflectors           (qr.c:99)  */
    float *YWt;   /* PR x PR  : (Y W^T), hoisted out of the trailing loop */
    float *tau;   /* PC                                                  */
    float *v;     /* PR       : current reflector, v[0] = 1              */
    float *z;     /* PR                                                  */
    float *acol;  /* PR                                                  */
} window_ws;

static int ws_alloc(window_ws *w, int PR, int PC)
{
    size_t pp = (size_t)PR * PC;
    w->panel = malloc(pp * sizeof(float));
    w->W = malloc(pp * sizeof(float));
    w->Y = malloc(pp * sizeof(float));
    w->YWt = malloc((size_t)PR * PR * sizeof(float));
    w->tau = malloc(PC * sizeof(float));
    w->v = malloc(PR * sizeof(float));
    w->z = malloc(PR * sizeof(float));
    w->acol = malloc(PR * sizeof(float));
    return w->panel && w->W && w->Y && w->YWt && w->tau && w->v && w->z && w->acol;
}

static void ws_free(window_ws *w)
{
    free(w->panel); free(w->W); free(w->Y); free(w->YWt);
    free(w->tau); free(w->v); free(w->z); free(w->acol);
}

/* Factor one PR x PC window held in ws->panel (qr.c:114-237). */
static void factor_window(window_ws *ws, int m, int PR, int PC, int pr, int pc)
{
    float *P = ws->panel, *W = ws->W, *Y = ws->Y, *v = ws->v, *z = ws->z;
    memset(W, 0, (size_t)PR * PC * sizeof(float));
    memset(Y, 0, (size_t)PR * PC * sizeof(float));
    for (int col = 0; col < PC; col++) {
        int vs, ve;
        reflector_span(m, PR, PC, pr, pc, col, &vs, &ve);
        int vlen = ve - vs;
        float *x = P + (size_t)col * PR;
        /* qr.c:144-152: norm, sign, u, tau -- sqrt is the double routine. */
        float ip = 0;
        for (int r = vs; r < ve; r++)
            ip += x[r] * x[r];
        float norm = sqrt(ip);
        float sign = (x[vs] < 0) ? -1.0 : 1.0;
        float u = x[vs] + sign * norm;
        float t = sign * u / norm;
        ws->tau[col] = t;
        x[vs] = -sign * norm;                      /* qr.c:158 */
        v[0] = 1;                                  /* qr.c:162-167 */
        for (int r = vs + 1; r < ve; r++) {
            x[r] /= u;
            v[r - vs] = x[r];
        }
        /* qr.c:170-202: z = -tau (v + W Y^T v), entry by entry. */
        for (int i = 0; i < PR; i++)
            z[i] = (i >= vs && i < ve) ? -t * v[i - vs] : 0;
        if (col > 0) {
            for (int i = 0; i < PR; i++) {
                float wytvi = 0;
                for (int j = vs; j < ve; j++) {
                    float wyt = 0;
                    for (int k = 0; k < col; k++)
                        wyt += W[(size_t)k * PR + i] * Y[(size_t)k * PR + j];
                    wytvi += wyt * v[j - vs];
                }
                z[i] -= t * wytvi;
            }
        }
        for (int i = 0; i < PR; i++)               /* qr.c:204-213 */
            W[(size_t)col * PR + i] = z[i];
        for (int i = 0; i < vlen; i++)
            Y[(size_t)col * PR + vs + i] = v[i];
        /* qr.c:215-235: apply H to the window's later columns; the reference
         * writes this as a dense (I - tau v v^T) row-times-column product. */
        for (int ac = col + 1; ac < PC; ac++) {
            float *a = P + (size_t)ac * PR;
            for (int i = 0; i < vlen; i++)
                ws->acol[i] = a[vs + i];
            for (int r = vs; r < ve; r++) {
                int vi = r - vs;
                float val = ws->acol[vi];
                for (int i = 0; i < vlen; i++)
                    val -= t * v[vi] * v[i] * ws->acol[i];
                a[r] = val;
            }
        }
    }
}

/*
 * qr.c:55-313.  In-place MMQR of the column-major m x n matrix `mat`
 * (lda = m).  `tau` must hold rowPanels*colPanels*PC floats (zero-filled
 * here, as qr.c:62 does); slot layout qr.c:300-304.
 * Returns 0, or -1 on allocation failure / illegal geometry arguments.
 */
ORACLE_API int oracle_mmqr(float *mat, float *tau, int m, int n, int PR, int PC)
{
    if (m < 1 || n < 1 || PR <= PC || PC < 1)
        return -1;
    int rowPanels, colPanels;
    oracle_getPanelDims(m, n, PR, PC, &rowPanels, &colPanels);
    memset(tau, 0, (size_t)rowPanels * colPanels * PC * sizeof(float));
    window_ws ws;
    if (!ws_alloc(&ws, PR, PC)) { ws_free(&ws); return -1; }
    int pcCount = 0;
    for (int pc = 0; pc < n; pc += PC, pcCount++) {
        int prCount = 0;
        for (int pr = m - PR; (pr + PR > pc) && pr >= 0; pr -= (PR - PC), prCount++) {
            for (int c = 0; c < PC; c++)           /* qr.c:81-87 */
                for (int r = 0; r < PR; r++)
                    ws.panel[(size_t)c * PR + r] = mat[(size_t)(c + pc) * m + r + pr];
            factor_window(&ws, m, PR, PC, pr, pc);
            for (int c = 0; c < PC; c++)           /* qr.c:242-248 */
                for (int r = 0; r < PR; r++)
                    mat[(size_t)(c + pc) * m + r + pr] = ws.panel[(size_t)c * PR + r];
            /* qr.c:255-293: A[pr:pr+PR, c] += (Y W^T) A[pr:pr+PR, c].  The
             * PR x PR product does not depend on c, so form it once with the
             * reference's own k-ordering. */
            if (pc + PC < n) {
                for (int i = 0; i < PR; i++)
                    for (int j = 0; j < PR; j++) {
                        float ywt = 0;
                        for (int k = 0; k < PC; k++)
                            ywt += ws.Y[(size_t)k * PR + i] * ws.W[(size_t)k * PR + j];
                        ws.YWt[(size_t)j * PR + i] = ywt;
                    }
                for (int c = pc + PC; c < n; c++) {
                    float *a = mat + (size_t)c * m + pr;
                    for (int i = 0; i < PR; i++)
                        ws.acol[i] = a[i];
                    for (int i = 0; i < PR; i++) {
                        float acc = 0;
                        for (int j = 0; j < PR; j++)
                            acc += ws.YWt[(size_t)j * PR + i] * ws.acol[j];
                        a[i] = ws.acol[i] + acc;
                    }
                }
            }
            for (int i = 0; i < PC; i++)           /* qr.c:300-304 */
                tau[(size_t)(rowPanels * pcCount + prCount) * PC + i] = ws.tau[i];
        }
    }
    ws_free(&ws);
    return 0;
}

/* qr.c:316-324 */
ORACLE_API void oracle_identity(float *A, int m)
{
    memset(A, 0, (size_t)m * m * sizeof(float));
    for (int i = 0; i < m; i++)
        A[(size_t)i * m + i] = 1;
}

/* qr.c:443-459: C(k x n) = A(k x m) B(m x n), column-major. */
ORACLE_API void oracle_dgemm(const float *A, const float *B, float *C, int k, int m, int n)
{
    for (int i = 0; i < n; i++)
        for (int j = 0; j < k; j++) {
            float c = 0;
            for (int l = 0; l < m; l++)
                c += A[(size_t)l * k + j] * B[(size_t)i * m + l];
            C[(size_t)i * k + j] = c;
        }
}

/*
 * qr.c:330-438.  R = triu(A) as m x n; Q = H_1 H_2 ... (generation order),
 * m x m.  The reference forms every dense H = I - tau v v^T (qr.c:415-423)
 * and multiplies Q*H with dgemm (qr.c:429).  Off the reflector's support H is
 * exactly the identity, so (Q H)(i, j) = Q(i, j) for columns j outside
 * [lo, hi) and a sum over l in [lo, hi) inside it; this restatement keeps
 * exactly those terms, in the reference's l-order, with H's entries rounded
 * the way qr.c:421 rounds them.
 */
ORACLE_API int oracle_explicitQR(const float *A, const float *tau, float *Q, float *R,
                                 int m, int n, int PR, int PC)
{
    for (int c = 0; c < n; c++)                    /* qr.c:334-343 */
        for (int r = 0; r < m; r++)
            R[(size_t)c * m + r] = (c >= r) ? A[(size_t)c * m + r] : 0;
    oracle_identity(Q, m);
    int rowPanels, colPanels;
    oracle_getPanelDims(m, n, PR, PC, &rowPanels, &colPanels);
    float *v = malloc(PR * sizeof(float));
    float *H = malloc((size_t)PR * PR * sizeof(float));
    float *qrow = malloc(PR * sizeof(float));
    if (!v || !H || !qrow) { free(v); free(H); free(qrow); return -1; }
    int pcCount = 0;
    for (int pc = 0; pc < n; pc += PC, pcCount++) {
        int prCount = 0;
        for (int pr = m - PR; (pr + PR > pc) && pr >= 0; pr -= (PR - PC), prCount++) {
            for (int col = 0; col < PC && col + pc < n; col++) {
                float t = tau[(size_t)(rowPanels * pcCount + prCount) * PC + col];
                int vs, ve;
                reflector_span(m, PR, PC, pr, pc, col, &vs, &ve);
                int lo = pr + vs, len = ve - vs;
                v[0] = 1;                          /* qr.c:397-407 */
                for (int i = 1; i < len; i++)
                    v[i] = A[(size_t)(pc + col) * m + lo + i];
                for (int j = 0; j < len; j++)      /* qr.c:415-423 */
                    for (int k = 0; k < len; k++) {
                        float h = (j == k) ? 1 : 0;
                        h -= t * v[k] * v[j];
                        H[(size_t)j * len + k] = h;
                    }
                for (int i = 0; i < m; i++) {      /* qr.c:426-429 */
                    for (int l = 0; l < len; l++)
                        qrow[l] = Q[(size_t)(lo + l) * m + i];
                    for (int j = 0; j < len; j++) {
                        float c = 0;
                        for (int l = 0; l < len; l++)
                            c += qrow[l] * H[(size_t)j * len + l];
                        Q[(size_t)(lo + j) * m + i] = c;
                    }
                }
            }
        }
    }
    free(v); free(H); free(qrow);
    return 0;
}

/* qr.c:468-474 input recipe: srand(seed); A[i] = (float)rand()/RAND_MAX in
 * linear column-major order.  glibc-specific stream. */
ORACLE_API void oracle_fill_rand(float *A, long count, unsigned seed)
{
    srand(seed);
    for (long i = 0; i < count; i++)
        A[i] = (float)rand() / RAND_MAX;
}
