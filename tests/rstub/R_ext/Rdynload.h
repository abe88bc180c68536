typedef struct _DllInfo DllInfo;
typedef struct { const char* name; DL_FUNC fun; int numArgs; } R_CallMethodDef;
int R_registerRoutines(DllInfo*, const void*, const R_CallMethodDef*, const void*, const void*);
int R_useDynamicSymbols(DllInfo*, int);
