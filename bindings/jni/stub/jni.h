/* DECLARATIONS-ONLY STAND-IN for the JDK's <jni.h>, so that calitas_b200_jni.c can be compile-checked in an image without a JDK
 * (tests/test_bindings.py).  It declares exactly the JNI types and JNIEnv entries the shim uses, with the signatures of the JNI
 * specification; it is not a JNI implementation and nothing built against it can be loaded by a JVM.  With a real JDK, build against
 * $JAVA_HOME/include instead and delete nothing here. */
#ifndef CALITAS_STUB_JNI_H
#define CALITAS_STUB_JNI_H
#include <stdint.h>
#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_ABORT 2
typedef int32_t jint; typedef int64_t jlong; typedef int8_t jbyte; typedef uint8_t jboolean; typedef jint jsize;
struct _jobject; typedef struct _jobject* jobject;
typedef jobject jclass; typedef jobject jstring; typedef jobject jarray; typedef jarray jobjectArray; typedef jarray jintArray; typedef jarray jlongArray; typedef jarray jbyteArray;
struct JNINativeInterface_; typedef const struct JNINativeInterface_* JNIEnv;
struct JNINativeInterface_ {
  jclass (*FindClass)(JNIEnv*, const char*);
  jint (*ThrowNew)(JNIEnv*, jclass, const char*);
  jsize (*GetArrayLength)(JNIEnv*, jarray);
  jobject (*GetObjectArrayElement)(JNIEnv*, jobjectArray, jsize);
  const char* (*GetStringUTFChars)(JNIEnv*, jstring, jboolean*);
  void (*ReleaseStringUTFChars)(JNIEnv*, jstring, const char*);
  void (*GetIntArrayRegion)(JNIEnv*, jintArray, jsize, jsize, jint*);
  void (*GetLongArrayRegion)(JNIEnv*, jlongArray, jsize, jsize, jlong*);
  void (*SetLongArrayRegion)(JNIEnv*, jlongArray, jsize, jsize, const jlong*);
  jbyte* (*GetByteArrayElements)(JNIEnv*, jbyteArray, jboolean*);
  void (*ReleaseByteArrayElements)(JNIEnv*, jbyteArray, jbyte*, jint);
  jobject (*NewDirectByteBuffer)(JNIEnv*, void*, jlong);
  void* (*GetDirectBufferAddress)(JNIEnv*, jobject);
  jlong (*GetDirectBufferCapacity)(JNIEnv*, jobject);
  void (*DeleteLocalRef)(JNIEnv*, jobject);
  void (*GetByteArrayRegion)(JNIEnv*, jbyteArray, jsize, jsize, jbyte*);
};
#endif
