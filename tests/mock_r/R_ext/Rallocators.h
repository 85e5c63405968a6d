/* mock of R_ext/Rallocators.h (R >= 3.0): custom allocator handed to Rf_allocVector3 */
#ifndef MOCK_RALLOCATORS_H
#define MOCK_RALLOCATORS_H
#include <stddef.h>
typedef struct R_allocator R_allocator_t;
typedef void* (*custom_alloc_t)(R_allocator_t* allocator, size_t);
typedef void (*custom_free_t)(R_allocator_t* allocator, void*);
struct R_allocator {
    custom_alloc_t mem_alloc; /* malloc equivalent */
    custom_free_t mem_free;   /* free equivalent */
    void* res;                /* reserved (maybe for copy) - must be NULL */
    void* data;               /* custom data for the allocator implementation */
};
#endif
