/* oracle/rstub/minc2.h -- TEST INFRASTRUCTURE.  In-memory stand-in for the
 * five libminc2 calls minc2_model makes (libminc/HDF5 are absent from this
 * image): a "file name" is looked up in a registry of double volumes that the
 * test harness filled (rstub_register_volume). */
#ifndef RSTUB_MINC2_H
#define RSTUB_MINC2_H
#ifdef __cplusplus
extern "C" {
#endif
typedef struct rstub_volume *mihandle_t;
typedef struct rstub_dim *midimhandle_t;
typedef unsigned long long misize_t;
#define MI_NOERROR 0
#define MI_ERROR (-1)
#define MI2_OPEN_READ 1
#define MI_TYPE_DOUBLE 6
#define MI_DIMCLASS_SPATIAL 1
#define MI_DIMATTR_ALL 0
#define MI_DIMORDER_FILE 0
int miopen_volume(const char *filename, int mode, mihandle_t *volume);
int miclose_volume(mihandle_t volume);
int miget_volume_dimensions(mihandle_t volume, int cls, int attr, int order,
                            int array_length, midimhandle_t dimensions[]);
int miget_dimension_sizes(const midimhandle_t dimensions[], misize_t array_length,
                          misize_t sizes[]);
int miget_real_value_hyperslab(mihandle_t volume, int buffer_data_type,
                               const misize_t start[], const misize_t count[],
                               void *buffer);
#ifdef __cplusplus
}
#endif
#endif
