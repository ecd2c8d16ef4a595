/* Stand-in for /root/reference/src/plume/hdf5calls.cc (the GH5_* seam declared
 * in /root/reference/src/plume/hdf5calls.h:33-59). HDF5 is not available in the
 * build image, so the reference's grid input (funkyNNNN.h5) and particle output
 * (pvplotNNN.h5part) are stored as flat record files instead:
 *
 *   record := u32 name_len | name bytes | i32 type | u64 count | u32 elem_size | payload
 *
 * type uses the reference's INTTYPE..CHARTYPE codes (hdf5calls.h:26-30); 100 is a
 * raw BucketStruct array. Group members are stored as "<group>/<dataset>".
 * Test infrastructure only: nothing in the product links this file. */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <hdf5.h>
#include <hdf5calls.h>

namespace {
struct OpenFile { FILE *fp; std::string path; char mode; };
struct OpenGroup { int file; std::string name; };
std::vector<OpenFile> g_files;
std::vector<OpenGroup> g_groups;
const long GROUP_BASE = 1L << 20;

FILE *file_of(hid_t id, std::string *prefix)
{
  if (id >= GROUP_BASE) {
    OpenGroup &g = g_groups[id - GROUP_BASE];
    if (prefix) *prefix = g.name + "/";
    return g_files[g.file].fp;
  }
  if (prefix) *prefix = "";
  return g_files[id - 1].fp;
}

size_t elem_size(int type)
{
  switch (type) {
    case INTTYPE: case UINTTYPE: case FLOATTYPE: return 4;
    case DOUBLETYPE: return 8;
    case CHARTYPE: return 1;
    case 100: return sizeof(BucketStruct);
  }
  fprintf(stderr, "gh5_flat: unknown type %d\n", type); exit(1);
}

void put_record(FILE *fp, const std::string &name, int type, unsigned long long count, const void *buf)
{
  unsigned nl = name.size(), es = elem_size(type);
  fseek(fp, 0, SEEK_END);
  fwrite(&nl, 4, 1, fp); fwrite(name.data(), 1, nl, fp);
  fwrite(&type, 4, 1, fp); fwrite(&count, 8, 1, fp); fwrite(&es, 4, 1, fp);
  fwrite(buf, es, count, fp);
}

/* find a record by name; leaves the stream at its payload */
bool find_record(FILE *fp, const std::string &name, int *type, unsigned long long *count, unsigned *es)
{
  fseek(fp, 0, SEEK_SET);
  for (;;) {
    unsigned nl;
    if (fread(&nl, 4, 1, fp) != 1) return false;
    std::string nm(nl, '\0');
    if (fread(&nm[0], 1, nl, fp) != nl) return false;
    if (fread(type, 4, 1, fp) != 1 || fread(count, 8, 1, fp) != 1 || fread(es, 4, 1, fp) != 1) return false;
    if (nm == name) return true;
    fseek(fp, (long) (*count * *es), SEEK_CUR);
  }
}

herr_t read_into(hid_t fp, const char *path, void *buf)
{
  int type; unsigned long long count; unsigned es;
  FILE *f = file_of(fp, NULL);
  if (!find_record(f, path, &type, &count, &es)) {
    fprintf(stderr, "gh5_flat: dataset %s not found\n", path); exit(1);
  }
  return fread(buf, es, count, f) == count ? 0 : -1;
}
}

hid_t GH5_fopen_serial(const char *filename, char mode)
{
  const char *m = mode == 'w' ? "w+b" : mode == 'a' ? "r+b" : "rb";
  FILE *fp = fopen(filename, m);
  if (!fp) {
    fprintf(stderr, "Unable to open %s. Make sure its in the current directory\n", filename);
    exit(0);
  }
  OpenFile of = { fp, filename, mode };
  g_files.push_back(of);
  return (hid_t) g_files.size();
}

hid_t GH5_fopen_parallel(const char *filename, char mode) { return GH5_fopen_serial(filename, mode); }

herr_t H5Fclose(hid_t id)
{
  OpenFile &of = g_files[id - 1];
  if (of.fp) { fclose(of.fp); of.fp = NULL; }
  return 0;
}
herr_t H5Gclose(hid_t) { return 0; }

int GH5_getsize(hid_t fp, const char *fullpath, hsize_t *dims)
{
  int type; unsigned long long count; unsigned es;
  if (!find_record(file_of(fp, NULL), fullpath, &type, &count, &es)) { dims[0] = 0; return -1; }
  dims[0] = count; dims[1] = 1;
  return 1;
}

hid_t GH5_gopen(hid_t fp, const char *name, char)
{
  OpenGroup g = { (int) (fp - 1), name };
  g_groups.push_back(g);
  return GROUP_BASE + (hid_t) g_groups.size() - 1;
}

herr_t GH5_readdata(hid_t fp, const char *p, double *b) { return read_into(fp, p, b); }
herr_t GH5_readdata(hid_t fp, const char *p, int *b) { return read_into(fp, p, b); }
herr_t GH5_readdata(hid_t fp, const char *p, unsigned *b) { return read_into(fp, p, b); }
herr_t GH5_read_grid_data(hid_t fp, const char *p, BucketStruct *c) { return read_into(fp, p, c); }

herr_t GH5_write_grid_data(hid_t fp, const char *path, int size, BucketStruct *c)
{
  std::string pre; FILE *f = file_of(fp, &pre);
  put_record(f, pre + path, 100, size, c);
  return 0;
}

herr_t GH5_WriteS(hid_t fp, const char *path, int *dims, void *buf, int, int, const int type)
{
  std::string pre; FILE *f = file_of(fp, &pre);
  unsigned long long n = (unsigned long long) dims[0] * (dims[1] > 1 ? dims[1] : 1);
  put_record(f, pre + path, type, n, buf);
  return 0;
}

herr_t GH5_WriteP(hid_t fp, const char *path, int *dims, void *buf, int a, int b, const int type)
{ return GH5_WriteS(fp, path, dims, buf, a, b, type); }
